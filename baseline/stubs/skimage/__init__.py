"""Stand-in for scikit-image (data/mtl_ds.py:33 of the reference: edge-label thinning in the dataset class)."""
