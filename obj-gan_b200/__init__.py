"""objgan_b200 -- Blackwell-native (sm_100a) hot path of Obj-GAN's image_generation
training step.  See DESIGN.md.  Import as ``objgan_b200`` (shim package at repo root)."""
__version__ = "0.1.0"
