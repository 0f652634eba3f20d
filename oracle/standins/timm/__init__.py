"""Stand-in for the `timm` package (TEST INFRASTRUCTURE ONLY).

The reference imports `PatchEmbed, Attention, Mlp` from
`timm.models.vision_transformer` (image_model/models.py:16) but neither vendors
nor pins timm, and timm is not installed in this image.  This stand-in restates
timm's published semantics (>= 0.9) for exactly those three classes so that
`oracle/make_golden.py` can import the unmodified reference from
/root/reference.  It is never imported by the product package.
"""
