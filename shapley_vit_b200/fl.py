"""Host-side mirror of the reference's federated_learning pieces that sit on the utility loop.

Same names, argument meaning and error behaviour as
  * ``ClientBase``                              reference federated_learning/client2.py:7-42
  * ``ServerBase`` (get_agg_ratio, model_agg_lazy, global_model, valid_loader)
                                                reference federated_learning/server2.py:15-127
  * ``get_difference_between_network_weights``  reference federated_learning/utils.py:735-749
  * ``get_aggregated_model``                    reference federated_learning/utils.py:781-792
  * ``evaluation``                              reference federated_learning/utils.py:864-926
but the arithmetic runs in libsvit on the GPU.  Nothing here falls back to torch operators.
"""
from __future__ import annotations

import copy
import math
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from .engine import CoalitionEngine, ValidationSet
from .models.vit import infer_config


def _state_dict_of(net) -> Dict[str, torch.Tensor]:
    return net if isinstance(net, dict) else net.state_dict()


def _clean_keys(sd: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """Drop DataParallel / PEFT wrappers' key prefixes (reference start.py:274-285)."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in sd.items():
        while k.startswith("module.") or k.startswith("base_model.model."):
            k = k[len("module."):] if k.startswith("module.") else k[len("base_model.model."):]
        out[k] = v
    return out


def config_of(net, heads: Optional[int] = None):
    cfg = getattr(net, "cfg", None)
    if cfg is not None:
        return cfg
    hf = getattr(net, "config", None)
    if hf is not None and hasattr(hf, "num_attention_heads"):
        heads = heads or hf.num_attention_heads
    return infer_config(_clean_keys(_state_dict_of(net)), heads=heads,
                        ln_eps=getattr(hf, "layer_norm_eps", 1e-12) if hf is not None else 1e-12)


def get_difference_between_network_weights(net_1, net_2) -> "OrderedDict[str, torch.Tensor]":
    """Delta[k] = W_1[k] - W_2[k] for every state_dict key (one-time preparation, one pass per
    client; elementwise fp32 subtraction on whatever device the tensors live on)."""
    sd1, sd2 = _state_dict_of(net_1), _state_dict_of(net_2)
    return OrderedDict((k, sd1[k] - sd2[k]) for k in sd1.keys())


def get_aggregated_model(nets: Sequence[Dict[str, torch.Tensor]], ratio: Sequence[float]):
    """sum_j ratio[j] * nets[j], key by key, through svit_aggregate (one launch for all keys).
    Returns None for an empty list, asserts len(nets) == len(ratio), like the reference."""
    if len(nets) == 0:
        return None
    assert len(nets) == len(ratio), f"len(nets)={len(nets)}, len(ratio)={len(ratio)}"
    from . import ops

    keys = list(nets[0].keys())
    dev = torch.device("cuda:0")
    flat = torch.stack([torch.cat([n[k].reshape(-1).to(torch.float32) for k in keys]) for n in nets])
    P = flat.shape[1]
    stride = (P + 7) // 8 * 8
    stacked = torch.zeros((len(nets), stride), dtype=torch.float32, device=dev)
    stacked[:, :P] = flat.to(dev)
    r = torch.tensor([list(ratio)], dtype=torch.float64).to(torch.float32)
    out = ops.aggregate(stacked, None, r, P=P)[0, :P]
    res: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    o = 0
    for k in keys:
        n = nets[0][k].numel()
        res[k] = out[o:o + n].reshape(nets[0][k].shape).to(nets[0][k].device)
        o += n
    return res


class _EvalCache:
    """Validation sets already resident on the device, keyed by the loader object.  The entry holds the loader
    itself (so its id cannot be recycled by another object while the entry lives) and the cache is bounded:
    the least recently used set is dropped, which frees its images in HBM."""
    MAX_SETS = 4
    sets: "OrderedDict[tuple, tuple]" = OrderedDict()

    @classmethod
    def get(cls, loader, cfg, precision: int, device) -> ValidationSet:
        key = (id(loader), precision, str(device), cfg)
        hit = cls.sets.get(key)
        if hit is not None and hit[0] is loader:
            cls.sets.move_to_end(key)
            return hit[1]
        vs = ValidationSet.from_loader(cfg, loader, precision, device)
        cls.sets[key] = (loader, vs)
        while len(cls.sets) > cls.MAX_SETS:
            cls.sets.popitem(last=False)
        return vs

    @classmethod
    def clear(cls) -> None:
        cls.sets.clear()


def evaluation(args, net, eval_loader):
    """(accuracy, mean cross-entropy) of ``net`` over ``eval_loader``; raises
    ``ValueError('loss is nan')`` exactly where the reference does.  ``args`` may carry the
    additive keys 'precision', 'device', 'image_chunk', 'heads'."""
    args = args if isinstance(args, dict) else {}
    precision = _lib.PRECISIONS[args.get("precision", _lib.DEFAULT_PRECISION)]
    device = args.get("device", "cuda:0")
    sd = _clean_keys(_state_dict_of(net))
    if any(".lora_A." in k for k in sd):   # PEFT-LoRA model (reference start.py:274-283): fold the factors in
        from . import lora
        from .models.vit import infer_config
        sd = lora.merged_state_dict(sd, args.get("lora_alpha", 8.0))
        cfg = getattr(net, "cfg", None) or infer_config(sd, heads=args.get("heads"))
    else:
        cfg = config_of(net, args.get("heads"))
    val = _EvalCache.get(eval_loader, cfg, precision, device)
    eng = CoalitionEngine(cfg, None, [sd], val, precision=precision, coalition_batch=1,
                          image_chunk=args.get("image_chunk", 128), device=device)
    correct, loss_sum = eng.evaluate([[1.0]])
    n = val.n
    if math.isnan(loss_sum[0]):
        raise ValueError("loss is nan")
    return correct[0] / n, loss_sum[0] / n


class ClientBase(object):
    """Data holder: only ``id`` and ``num_local_data_train = len(train_set)`` matter here."""

    def __init__(self, id, args, net_train, train_set, test_set=None):
        self.id = id
        self.args = args
        self.local_data_train = train_set
        self.num_local_data_train = len(self.local_data_train)
        if test_set is not None:
            self.local_data_test = test_set
            self.num_local_data_test = len(self.local_data_test)
        self.optimizer = None


class ServerBase(object):
    """Holds the global model, the clients and the validation loader."""

    def __init__(self, args, net_train, clients, test_set, valid_set=None, group_valid_dataset=None):
        self.args = args
        self.global_model = copy.deepcopy(net_train)
        self.clients = clients
        self.num_clients = len(self.clients)
        self.valid_loader = valid_set
        self.group_valid_loader = []

    @property
    def global_model_state(self):
        return copy.deepcopy(_state_dict_of(self.global_model))

    def get_agg_ratio(self, selected_clients=None) -> List[float]:
        """FedAvg coefficients n_j / sum_k n_k over the given clients (Python floats)."""
        if selected_clients is None:
            selected_clients = self.clients
        n_train_list = [client.num_local_data_train for client in selected_clients]
        total = sum(n_train_list)
        return [n / total for n in n_train_list]

    def model_agg_lazy(self, init_global_model, client_models):
        """global_model <- W_0 + sum_t client_models[t] (a list of per-round aggregates)."""
        w = OrderedDict((k, v.clone()) for k, v in _state_dict_of(init_global_model).items())
        for agg in client_models:
            for key in agg.keys():
                w[key] = w[key] + agg[key].to(w[key].device)
        if isinstance(self.global_model, dict):
            self.global_model = w
        else:
            self.global_model.load_state_dict(w)
