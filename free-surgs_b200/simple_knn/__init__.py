"""PyTorch-on-device stand-in for the reference's ``simple_knn`` extension (submodules/simple-knn)."""
