"""Stand-in for timm==0.9.2 (requirements.txt:17 of the reference; not installed in this image): only the three
symbols the reference imports from `timm.models.layers` (models/swin_transformer_mtlora.py:19)."""
