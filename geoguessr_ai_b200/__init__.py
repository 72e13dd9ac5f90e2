"""B200-native (sm_100a) implementation of geoguessr-ai's post-encoder geolocation path.

    from geoguessr_ai_b200 import SuperGuessr, ProtoRefiner, ModelOutput

Same module contract as the reference's ``models.super_guessr.SuperGuessr`` and
``models.proto_refiner.ProtoRefiner``; the arithmetic runs in libgeoguessr_b200.so (hand-written
CUDA: TMA + tcgen05/TMEM GEMMs, fused bandwidth-bound kernels) behind the C ABI of
``include/geoguessr_b200.h``.  No Triton, no torch.compile, no CPU fallback.
"""
from .utils import ModelOutput, TopK  # noqa: F401
from .super_guessr import SuperGuessr, dp_chunk_bounds  # noqa: F401
from .proto_refiner import ProtoRefiner, shard_cells  # noqa: F401
from . import embedding_store, ops, proto_builder, synth  # noqa: F401

__all__ = ["SuperGuessr", "ProtoRefiner", "ModelOutput", "TopK", "ops", "synth", "shard_cells", "dp_chunk_bounds",
           "embedding_store", "proto_builder"]
