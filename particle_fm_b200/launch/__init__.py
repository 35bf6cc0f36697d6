"""Multi-GPU launchers of the hot path: one process per GPU (torchrun), torch.distributed for the plumbing.

generate.py  jets sharded over the ranks, no data-path collective, one final gather   (generation)
train.py     data-parallel training: all-reduce of the flat folded-weight gradient    (training)
"""
from .generate import block_noise, generate_data_sharded, integrate_and_gather, shard_bounds  # noqa: F401
from .train import GraphedTrainStep, attach_flat_grad_allreduce, detach_flat_grad_allreduce  # noqa: F401
