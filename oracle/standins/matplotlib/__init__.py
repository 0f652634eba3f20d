"""Stand-in for matplotlib (TEST INFRASTRUCTURE ONLY): the reference only uses
`pyplot.imsave` as a per-step debugging side effect
(image_model/diffusion/gaussian_diffusion.py:16,796)."""
