"""Drop-in package name.  ``scene/gaussian_model.py:20`` of the reference does
``from simple_knn._C import distCUDA2``; with this repository on ``sys.path`` that import
resolves to ``simple_knn/_C.py`` here and runs on the sm_100a kernels."""
