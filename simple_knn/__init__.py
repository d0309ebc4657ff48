"""Drop-in module name: SCGaussian does `from simple_knn._C import distCUDA2`
(reference scene/gaussian_model.py:20).  With this repo on PYTHONPATH that import resolves to the
B200 library (scgr_knn3_mean_dist2, scgaussian_b200/csrc/knn.cu)."""
