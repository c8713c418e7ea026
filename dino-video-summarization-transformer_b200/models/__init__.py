"""Mirror of the reference's `models` package import surface (models/__init__.py): `from models import
get_vit_base_patch16_224` (dino_similarity.py:11, eval_knn.py:25, eval_linear.py:24, train_ssl.py:40)."""
from .timesformer import (TimeSformer, VisionTransformer, get_vit_base_patch16_224,  # noqa: F401
                          vit_base_patch16_224)
