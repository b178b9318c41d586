"""Drop-in replacement for the reference's `Models` module.

`Model_Pretraining.py` does `from Models import HSIMAE` and `Model_Finetuning.py`
does `from Models import DualViT, HSIViT` (/root/reference/Model_Pretraining.py:13,
Model_Finetuning.py:14); with this repository ahead of the reference on
``sys.path`` those imports resolve here and the unchanged drivers run on the
sm_100a CUDA path.
"""
from hsimae_b200.modules import (HSIMAE, DualViT, HSIViT, PatchEmbed, Attention, SwiGLU, Block,  # noqa: F401
                                 DropPath)
from hsimae_b200.host import sincos_table as get_3d_sincos_pos_embed_table  # noqa: F401

__all__ = ["HSIMAE", "DualViT", "HSIViT"]
