"""``simple_knn._C`` of the reference exports exactly one function (submodules/simple-knn/ext.cpp:14-16)."""
from binocular3dgs_b200.simple_knn import distCUDA2  # noqa: F401
