#!/usr/bin/env bash
# baseline/build_reference.sh — install the UNMODIFIED reference rasterizer for sm_100a.
#
# Output goes only into baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to
# the GPU box, where /root/reference does not exist):
#
#   baseline/_ref/diff_gaussian_rasterization/{__init__.py,_C*.so}   the reference's stock
#       package, built by its own setup.py (rasterize_points.cu + ext.cpp + cuda_rasterizer/)
#   baseline/_ref/reference_tree/{gaussian_renderer,scene,utils,arguments}/   the reference's
#       Python files that call the operator (render(), GaussianModel), byte-for-byte, so
#       tests/test_gpu_render_adapter.py can execute the reference's real render() on
#       both operators on the GPU box
#   baseline/_ref/BUILD_INFO.json   what was built, from where, with which toolchain
#
# The reference tree is read-only, and its setup.py builds in-tree, so the install runs
# from a scratch copy under /tmp.  No file of the reference is modified.
set -euo pipefail
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
OUT="$ROOT/baseline/_ref"
if [ ! -d "$REF/submodules/diff-gaussian-rasterization" ]; then
  echo "build_reference.sh: $REF not present (GPU box?) — using the prebuilt baseline/_ref" >&2
  exit 0
fi
STAMP="$OUT/BUILD_INFO.json"
if [ -f "$STAMP" ] && ls "$OUT"/diff_gaussian_rasterization/_C*.so >/dev/null 2>&1 && [ -z "${FORCE:-}" ]; then
  echo "build_reference.sh: baseline/_ref already built (FORCE=1 to rebuild)"; exit 0
fi
TMP="$(mktemp -d /tmp/dgr_ref.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF/submodules/diff-gaussian-rasterization" "$TMP/dgr"
mkdir -p "$OUT"
T0=$(date +%s)
( cd "$TMP" && TORCH_CUDA_ARCH_LIST="10.0a" MAX_JOBS="${MAX_JOBS:-8}" \
  python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
         --upgrade --target "$OUT" "$TMP/dgr" )
T1=$(date +%s)
mkdir -p "$OUT/reference_tree"
for d in gaussian_renderer scene utils arguments; do
  [ -e "$OUT/reference_tree/$d" ] && chmod -R u+w "$OUT/reference_tree/$d" && rm -rf "$OUT/reference_tree/$d"
  cp -r "$REF/$d" "$OUT/reference_tree/$d"
  chmod -R u+w "$OUT/reference_tree/$d"
done
python - "$OUT" "$REF" "$((T1-T0))" <<'PY'
import hashlib, json, os, subprocess, sys
out, ref, secs = sys.argv[1], sys.argv[2], int(sys.argv[3])
so = [f for f in os.listdir(os.path.join(out, "diff_gaussian_rasterization")) if f.startswith("_C")][0]
def sha(p):
    return hashlib.sha256(open(p, "rb").read()).hexdigest()[:16]
src = os.path.join(ref, "submodules", "diff-gaussian-rasterization")
files = ["setup.py", "ext.cpp", "rasterize_points.cu", "rasterize_points.h", "cuda_rasterizer/forward.cu",
         "cuda_rasterizer/backward.cu", "cuda_rasterizer/rasterizer_impl.cu", "diff_gaussian_rasterization/__init__.py"]
import torch
info = {"what": "reference diff_gaussian_rasterization, stock setup.py, unmodified sources",
        "arch": "TORCH_CUDA_ARCH_LIST=10.0a", "build_seconds": secs, "extension": so,
        "extension_sha256_16": sha(os.path.join(out, "diff_gaussian_rasterization", so)),
        "source_sha256_16": {f: sha(os.path.join(src, f)) for f in files},
        "torch": torch.__version__,
        "nvcc": subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]}
json.dump(info, open(os.path.join(out, "BUILD_INFO.json"), "w"), indent=1)
print(json.dumps(info))
PY
