"""Stand-in for the absent `matplotlib`: inference_ddp.py:31 imports pyplot and never draws with it."""
