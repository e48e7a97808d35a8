"""shapley-vit on B200: the client-contribution utility loop (FedAvg coalition aggregation +
batched ViT scoring + Shapley accumulation) on hand-written sm_100a kernels behind a C ABI.

The package holds only what that path needs:
  csrc/        CUDA kernels + the C ABI (libsvit.so, declared in include/svit.h)
  _lib, ops    ctypes binding and the tensor-level front end (no fallback: CUDA or error)
  layout       ViT geometry, HF state_dict order, plan layout
  engine       device-resident game state, batched coalition evaluation
  game, fl     mirrors of the reference's Game / ServerBase / ClientBase / evaluation
  estimators, compared   the Shapley estimators (plan -> prefetch -> accumulate)
  dist         coalition sharding across GPUs (torch.distributed)
  synth        deterministic synthetic inputs
"""
__version__ = "0.1.0"
