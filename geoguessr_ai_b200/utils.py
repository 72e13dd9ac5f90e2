"""Return types of the drop-in modules (reference: models/utils.py:12-17)."""
from collections import namedtuple

# same field names and order as the reference's ModelOutput
ModelOutput = namedtuple("ModelOutput", "loss loss_clf preds_LLH preds_geocell top5_geocells embedding")

# torch.return_types.topk look-alike: attribute access (.values / .indices) and tuple unpacking,
# both of which the reference's callers use (inference.py:173-174, main_coordinator_idun_s3.py:400-403)
TopK = namedtuple("topk", ["values", "indices"])

LABEL_SMOOTHING_CONSTANT = 65  # config.py:52 (PIGEOTTO)
CLIP_EMBED_DIM = 1024  # config.py
CLIP_PRETRAINED_HEAD = "saved_models/New_Base_smooth_avg_MT_Geo_SV.model"  # config.py:60
