"""B200-native (sm_100a) implementation of the JPDVT hot path: denoiser, diffusion step, position-to-grid assignment.

Public surface mirrors the reference's own modules:
    jpdvt_mt_ntnu_b200.models     <-> image_model/models.py
    jpdvt_mt_ntnu_b200.diffusion  <-> image_model/diffusion/
    jpdvt_mt_ntnu_b200.assignment <-> the find_permutation / pairwise_distances snippet of image_model/inference.py
All arithmetic runs in libjpdvt_sm100.so (include/jpdvt_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
