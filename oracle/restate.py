"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's utility loop.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
(``shapley_vit_b200``) never does; it fails loudly without its CUDA library.

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), and the ViT arithmetic lives in the unpinned third-party
``transformers`` package (``start.py:19, 258-267``; this image has 5.5.0).  This
restatement is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, run in
the build container through ``oracle/ref_shim.py`` with HF's
``ViTForImageClassification`` as the model; ``oracle/make_golden.py`` is the
generating script and ``tests/golden/*.npz|json`` the committed fixtures
(``tests/test_oracle_golden.py`` checks this file against them).

One exception, said where it applies: the LoRA branch (``split_peft_state_dict`` and
``vit_forward(..., lora=...)``) restates PEFT's published
``lora.Linear`` forward -- ``peft`` is not installed in the build container, so for
that row the parity is UNPINNED (the entry-wise aggregation underneath it is the
reference's own, pinned as above).

Each function cites the reference lines it follows.  All arithmetic is fp32 on
CPU tensors, in the reference's operation order.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Iterable, List, Sequence, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# F1-F4: weight differences, FedAvg ratios, aggregation
# --------------------------------------------------------------------------- #

def get_difference_between_network_weights(sd_1: Dict[str, torch.Tensor],
                                           sd_2: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """Delta[k] = W_1[k] - W_2[k] per state_dict key.
    Follows reference federated_learning/utils.py:735-749."""
    return OrderedDict((k, sd_1[k] - sd_2[k]) for k in sd_1.keys())


def get_agg_ratio(n_train_list: Sequence[int]) -> List[float]:
    """r_j = n_j / sum(n) over the coalition members only (Python floats).
    Follows reference federated_learning/server2.py:68-81."""
    total = sum(n_train_list)
    return [n / total for n in n_train_list]


def get_aggregated_model(nets: Sequence[Dict[str, torch.Tensor]], ratio: Sequence[float]):
    """agg[k] = r_0*D_0[k]; agg[k] = agg[k] + r_i*D_i[k]  (product and sum rounded
    separately, left to right).  Follows reference federated_learning/utils.py:781-792."""
    if len(nets) == 0:
        return None
    assert len(nets) == len(ratio)
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for i, w in enumerate(nets):
        for key in w.keys():
            if i == 0:
                out[key] = ratio[i] * w[key]
            else:
                out[key] = out[key] + ratio[i] * w[key]
    return out


def model_agg_lazy(w0: Dict[str, torch.Tensor], per_round_aggregates: Sequence[Dict[str, torch.Tensor]]):
    """W_S[k] = W_0[k] + sum_t agg_t[k].  Follows reference
    federated_learning/server2.py:121-127 (the load_state_dict is the caller's)."""
    w = OrderedDict((k, v.clone()) for k, v in w0.items())
    for agg in per_round_aggregates:
        for key in agg.keys():
            w[key] = w[key] + agg[key]
    return w


def coalition_state_dict(w0, deltas: Sequence[Dict[str, torch.Tensor]], n_train: Sequence[int],
                         members_in_order: Sequence[int]):
    """Game.eval_utility's model construction (reference
    fed_client_contribution/game.py:88-100) for an explicit member order."""
    members = list(members_in_order)
    ratio = get_agg_ratio([n_train[j] for j in members])
    agg = get_aggregated_model([deltas[j] for j in members], ratio)
    return model_agg_lazy(w0, [agg] if agg is not None else [])


def lazy_state_dict(w0, round_deltas, selection_matrix, n_train: Sequence[int], coalition: Sequence[int],
                    current_round: int, include_from_round: int = 0):
    """The model compute_utilities_lazy reconstructs for one subset (reference
    fed_client_contribution/utils_fed_shapley.py:168-183): for every round t in
    [include_from_round, current_round] the FedAvg aggregate of the members selected in that round,
    added to W_0 in round order (model_agg_lazy, server2.py:121-127)."""
    per_round = []
    for t in range(current_round + 1):
        if t < include_from_round:
            continue
        members = [j for j in coalition if selection_matrix[t][j]]
        if members:
            per_round.append(get_aggregated_model([round_deltas[t][j] for j in members],
                                                  get_agg_ratio([n_train[j] for j in members])))
    return model_agg_lazy(w0, per_round)


def compute_utilities_lazy(w0, round_deltas, selection_matrix, n_train, cfg, images, labels, previous_utility,
                           current_round: int, include_from_round: int = 0):
    """Multi-round "lazy" utilities of every NON-EMPTY subset in powerset order.
    Follows reference fed_client_contribution/utils_fed_shapley.py:146-196.
    Returns (utilities [2][n_subsets], subsets)."""
    n = len(n_train)
    subsets = list(powerset(range(n)))
    acc, loss = [], []
    for S in subsets:
        sd = lazy_state_dict(w0, round_deltas, selection_matrix, n_train, S, current_round, include_from_round)
        a, l = evaluation(sd, cfg, images, labels)
        acc.append(a - previous_utility[0])
        loss.append(l - previous_utility[1])
    return [acc, loss], subsets


def reference_member_order(coalition: Iterable[int], selection: Sequence[bool] | None = None) -> List[int]:
    """The order in which the reference sums a coalition: iteration order of
    ``frozenset(coalition)`` filtered by the selection vector
    (reference game.py:75, 90-91).  Usually ascending, not always."""
    fs = frozenset(int(j) for j in coalition)
    return [j for j in fs if selection is None or selection[j]]


# --------------------------------------------------------------------------- #
# M1: HF ViTForImageClassification forward, restated on a raw state_dict
# --------------------------------------------------------------------------- #

def split_peft_state_dict(sd: Dict[str, torch.Tensor]):
    """PEFT-LoRA state_dict (reference start.py:274-283: r=16, alpha=8 on query/value, classifier in
    modules_to_save; DataParallel ``module.`` prefix) -> (HF-keyed base state_dict, {(layer, proj): (A, B)}).
    PEFT is absent from the build container; the key scheme is the published one of peft.tuners.lora
    (``base_model.model.<path>.{base_layer.weight, lora_A.<adapter>.weight, lora_B.<adapter>.weight}``,
    ``classifier.{original_module, modules_to_save.<adapter>}``).  PARITY UNPINNED for this function."""
    import re

    hf, lora = {}, {}
    for k, v in sd.items():
        while k.startswith("module.") or k.startswith("base_model.model."):
            k = k.split(".", 1)[1] if k.startswith("module.") else k[len("base_model.model."):]
        m = re.match(r"^vit\.encoder\.layer\.(\d+)\.attention\.attention\.(\w+)\.lora_([AB])\.[^.]+\.weight$", k)
        if m:
            a_b = lora.setdefault((int(m.group(1)), m.group(2)), [None, None])
            a_b[0 if m.group(3) == "A" else 1] = v
            continue
        if k.startswith("classifier.original_module."):
            continue
        k = re.sub(r"^classifier\.modules_to_save\.[^.]+\.", "classifier.", k.replace(".base_layer.", "."))
        hf[k] = v
    return hf, {key: (a, b) for key, (a, b) in lora.items()}


def vit_forward(sd: Dict[str, torch.Tensor], cfg, images: torch.Tensor,
                return_hidden: bool = False, lora=None, lora_scaling: float = 0.5):
    """logits = ViTForImageClassification(images).logits with weights ``sd``.

    Restates transformers 5.5.0 models/vit/modeling_vit.py: embeddings :100-128
    (16x16/s16 conv -> flatten -> [CLS] concat -> +pos), self-attention :171-196,
    :220-252 (softmax(QK^T * d^-0.5) V, no mask), layer :328-347 (pre-LN,
    residuals in the layer), exact-erf GELU, final LayerNorm, CLS -> Linear
    :620-654.  ``cfg`` is a shapley_vit_b200.layout.VitConfig (only geometry)."""
    h, nh, hd = cfg.hidden, cfg.heads, cfg.head_dim
    B = images.shape[0]
    e = "vit.embeddings."
    x = F.conv2d(images, sd[e + "patch_embeddings.projection.weight"],
                 sd[e + "patch_embeddings.projection.bias"], stride=cfg.patch)
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat((sd[e + "cls_token"].expand(B, -1, -1), x), dim=1)
    x = x + sd[e + "position_embeddings"]
    hidden = [x] if return_hidden else None
    scaling = hd ** -0.5
    for i in range(cfg.layers):
        p = f"vit.encoder.layer.{i}."
        y = F.layer_norm(x, (h,), sd[p + "layernorm_before.weight"], sd[p + "layernorm_before.bias"], cfg.ln_eps)
        q = F.linear(y, sd[p + "attention.attention.query.weight"], sd[p + "attention.attention.query.bias"])
        k = F.linear(y, sd[p + "attention.attention.key.weight"], sd[p + "attention.attention.key.bias"])
        v = F.linear(y, sd[p + "attention.attention.value.weight"], sd[p + "attention.attention.value.bias"])
        if lora:  # peft lora.Linear.forward in eval mode: base(x) + lora_B(lora_A(x)) * (alpha / r), unmerged
            if (i, "query") in lora:
                q = q + F.linear(F.linear(y, lora[(i, "query")][0]), lora[(i, "query")][1]) * lora_scaling
            if (i, "value") in lora:
                v = v + F.linear(F.linear(y, lora[(i, "value")][0]), lora[(i, "value")][1]) * lora_scaling
        q, k, v = (t.view(B, -1, nh, hd).transpose(1, 2) for t in (q, k, v))
        att = torch.softmax(torch.matmul(q, k.transpose(2, 3)) * scaling, dim=-1)
        ctx = torch.matmul(att, v).transpose(1, 2).reshape(B, -1, h)
        x = F.linear(ctx, sd[p + "attention.output.dense.weight"], sd[p + "attention.output.dense.bias"]) + x
        y = F.layer_norm(x, (h,), sd[p + "layernorm_after.weight"], sd[p + "layernorm_after.bias"], cfg.ln_eps)
        y = F.gelu(F.linear(y, sd[p + "intermediate.dense.weight"], sd[p + "intermediate.dense.bias"]))
        x = F.linear(y, sd[p + "output.dense.weight"], sd[p + "output.dense.bias"]) + x
        if return_hidden:
            hidden.append(x)
    x = F.layer_norm(x, (h,), sd["vit.layernorm.weight"], sd["vit.layernorm.bias"], cfg.ln_eps)
    logits = F.linear(x[:, 0, :], sd["classifier.weight"], sd["classifier.bias"])
    return (logits, hidden) if return_hidden else logits


# --------------------------------------------------------------------------- #
# F5: evaluation
# --------------------------------------------------------------------------- #

@torch.no_grad()
def evaluation(sd, cfg, images: torch.Tensor, labels: torch.Tensor, batch_size: int = 128,
               return_details: bool = False):
    """Batches of ``batch_size`` (start.py:84 uses 128, no shuffle): logits, first-index
    argmax, integer correct count, fp32 CrossEntropy(sum) per batch accumulated
    into a Python float; NaN loss raises ``ValueError('loss is nan')``; returns
    (correct/n, loss/n).  Follows reference federated_learning/utils.py:864-926.
    (The reference leaves autograd on; results are identical under no_grad --
    SURVEY.md section 8(c)(10).)"""
    n = images.shape[0]
    correct, loss = 0, 0.0
    preds, all_logits = [], []
    for s in range(0, n, batch_size):
        img, lab = images[s:s + batch_size], labels[s:s + batch_size].long()
        out = vit_forward(sd, cfg, img)
        pred = out.argmax(dim=1)
        correct += pred.eq(lab).sum().item()
        loss += F.cross_entropy(out, lab, reduction="sum").item()
        if return_details:
            preds.append(pred)
            all_logits.append(out)
    if math.isnan(loss):
        raise ValueError("loss is nan")
    acc, loss_mean = correct / n, loss / n
    if return_details:
        return acc, loss_mean, dict(correct=correct, loss_sum=loss, pred=torch.cat(preds),
                                    logits=torch.cat(all_logits))
    return acc, loss_mean


# --------------------------------------------------------------------------- #
# F6: the memoised game
# --------------------------------------------------------------------------- #

class OracleGame:
    """Restates reference fed_client_contribution/game.py:6-36, 73-114 on raw
    state_dicts: memoised v(S) = [acc(S) - acc_0, loss(S) - loss_0], v({}) = [0, 0]."""

    def __init__(self, cfg, w0, deltas, n_train, images, labels, selection=None,
                 previous_utility=None, batch_size: int = 128):
        self.cfg, self.w0, self.deltas, self.n_train = cfg, w0, list(deltas), list(n_train)
        self.images, self.labels, self.batch_size = images, labels, batch_size
        self._n_all = len(self.deltas)
        self.client_selection_vector = list(selection) if selection is not None else [True] * self._n_all
        self.selected_clients = [i for i in range(self._n_all) if self.client_selection_vector[i]]
        self.n = len(self.selected_clients)
        self.utility_dim = 2
        if previous_utility is None:
            previous_utility = list(evaluation(w0, cfg, images, labels, batch_size))
        self.previous_utility = list(previous_utility)
        self.utility = [{}, {}]
        self.details = {}
        self.default_shapley_value = [{c: 0 for c in range(self._n_all)} for _ in range(2)]

    def eval_utility(self, coalition):
        coalition = frozenset(int(j) for j in coalition)
        if len(coalition) == 0:
            return [0, 0]
        if coalition in self.utility[0]:
            return [self.utility[i][coalition] for i in range(2)]
        members = [j for j in coalition if self.client_selection_vector[j]]
        sd = coalition_state_dict(self.w0, self.deltas, self.n_train, members)
        acc, loss, det = evaluation(sd, self.cfg, self.images, self.labels, self.batch_size,
                                    return_details=True)
        self.details[coalition] = det
        self.utility[0][coalition] = acc - self.previous_utility[0]
        self.utility[1][coalition] = loss - self.previous_utility[1]
        return [self.utility[i][coalition] for i in range(2)]


# --------------------------------------------------------------------------- #
# F7: exact Shapley (the one estimator the smoke test and bench need here; the
# product's estimators are pinned directly against reference outputs in
# tests/golden/estimators.json)
# --------------------------------------------------------------------------- #

def powerset(iterable):
    """Non-empty subsets by size then lexicographic; reference utils_shapley.py:141-144."""
    from itertools import chain, combinations

    s = list(iterable)
    it = chain.from_iterable(combinations(s, r) for r in range(1, len(s) + 1))
    return {tuple(sorted(t)): i for i, t in enumerate(it)}


def shapley_exact(game):
    """phi_i = sum_{S contains i} c(|S|-1) v(S) - sum_{S without i} c(|S|) v(S),
    c(s) = s!(n-s-1)!/n!.  Follows reference utils_shapley.py:185-203."""
    players, n = game.selected_clients, game.n
    sv = [{c: 0 for c in range(game._n_all)} for _ in range(2)]
    fact = math.factorial
    coef = {s: fact(s) * fact(n - s - 1) / fact(n) for s in range(n)}
    for S in list(powerset(players)):
        u = game.eval_utility(S)
        for i in range(2):
            for j in S:
                sv[i][j] += coef[len(S) - 1] * u[i]
            for j in set(players) - set(S):
                sv[i][j] -= coef[len(S)] * u[i]
    return sv
