"""hsimae_b200 -- B200 (sm_100a) native compute path for HSIMAE pretraining, dual-branch
fine-tuning and encoder-only classification, behind the reference's `Models.py` surface."""
from .modules import HSIMAE, DualViT, HSIViT  # noqa: F401

__all__ = ["HSIMAE", "DualViT", "HSIViT"]
