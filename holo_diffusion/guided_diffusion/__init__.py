"""Enums of the reference's guided_diffusion package that its configs name (the sampler itself lives in
holo_diffusion_b200/diffusion.py)."""
