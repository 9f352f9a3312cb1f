"""Golden cases: the same table tests/golden/make_golden.py generated the fixtures from."""
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden_cases", os.path.join(_HERE, "golden", "make_golden.py"))


def _load_cases():
    # parse CASES out of make_golden.py without importing torch-on-GPU machinery
    src = open(os.path.join(_HERE, "golden", "make_golden.py")).read()
    start = src.index("CASES = {")
    end = src.index("\n}\n", start) + 3
    ns = {}
    exec(src[start:end], {"dict": dict}, ns)
    return ns["CASES"]


CASES = _load_cases()


def load_golden(name):
    return dict(np.load(os.path.join(_HERE, "golden", name + ".npz")))


def make_case(name):
    from workloads import make_camera, make_pixel_grads, make_scene
    spec = CASES[name]
    scene = make_scene(**spec["scene"])
    cam = make_camera(**spec["cam"])
    grads = make_pixel_grads(cam.image_width, cam.image_height, seed=spec["scene"]["seed"] + 100)
    bg = np.asarray(spec["bg"], np.float32)
    return scene, cam, bg, grads, spec.get("scale_modifier", 1.0)
