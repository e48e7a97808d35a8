"""Entry point of the client-contribution valuation: drop-in for reference
shapleyserver/start.py (``start`` :248-303, ``getInitialShapleyValue`` :82-196,
``checkLocalTrainingModelExist`` :198-222), running on the batched sm_100a utility loop.

Flow (same as the reference): build the validation loader (batch 128, no shuffle), score the
initial global model -> previous_utility, load the client checkpoints
(``shapleyserver/local_training/client_<i>_model/ViT_epoch_9.pth.tar``, ``{'state_dict': ...}``),
form the per-client weight deltas, build ClientBase/ServerBase/Game and call the estimator.

Deliberate differences from the reference: the models are distinct
objects (the reference binds four names to one module, so every delta is zero), the number of
clients is not pinned to 3, the estimator's result is kept (``LAST_SHAPLEY_VALUE`` and the return
value of ``start()``) instead of only printed, and ``--synthetic`` replaces the private OCT data
and the pretrained HF checkpoint, neither of which ships with the reference (``--synthetic_data`` replaces the
data only: the client models are then read from the reference's checkpoint files, waiting for each to appear).
"""
from __future__ import annotations

import copy
import errno
import os
import time

import torch
from torch.utils.data import DataLoader

from .fed_client_contribution.game2 import Game
from .fed_client_contribution.utils_shapley import call_shapley_computation_method
from .federated_learning.client2 import ClientBase
from .federated_learning.server2 import ServerBase
from .federated_learning.utils import evaluation, get_difference_between_network_weights
from .models.vit import ViTForImageClassification
from .opts import opt, opts

try:  # the reference reads .env through python-dotenv (start.py:22-27)
    from dotenv import load_dotenv

    load_dotenv()
except Exception:  # pragma: no cover - optional dependency
    pass

my_local_model_path = os.getenv("LOCAL_MODEL_PATH")
my_global_model_path = os.getenv("GLOBAL_MODEL_PATH")
my_validation_dataset = os.getenv("VALIDATION_DATASET")

LAST_SHAPLEY_VALUE = None


def _loop_args():
    return {"precision": opt.dtype, "coalition_batch": opt.coalition_batch, "image_chunk": opt.image_chunk,
            "approximation_method": opt.approximation_method, "seed": opt.seed, "lora_alpha": opt.lora_alpha,
            **({"m": opt.mc_samples} if opt.mc_samples else {})}


def client_checkpoint_paths(num_clients: int):
    base = os.path.join(os.getcwd(), "shapleyserver", "local_training")
    return [os.path.join(base, f"client_{i + 1}_model", "ViT_epoch_9.pth.tar") for i in range(num_clients)]


def checkLocalTrainingModelExist(filepath, poll_seconds: float = 1.0):
    """Block until ``filepath`` exists and is not locked by a writer (reference :198-222)."""
    def is_file_locked(path):
        try:
            fd = os.open(path, os.O_RDWR | os.O_EXCL)
            os.close(fd)
            return False
        except OSError as e:
            return e.errno != errno.EEXIST

    while not (os.path.exists(filepath) and not is_file_locked(filepath)):
        print("Waiting for the file to be unlocked...")
        time.sleep(poll_seconds)
    return True


def getInitialShapleyValue(dataset, init_global_model, *client_models, client_sizes=None, checkpoints=None,
                           args=None):
    """Returns ``(shapley_value_all_rounds, shapley_value_sum)`` -- the equal-split initial values,
    exactly what the reference returns; the estimator's vector is left in LAST_SHAPLEY_VALUE."""
    global LAST_SHAPLEY_VALUE
    args = dict(_loop_args() if args is None else args)
    valid_loader = DataLoader(dataset, batch_size=128, shuffle=False)
    fed_valid_acc, fed_valid_loss = evaluation(args, init_global_model, valid_loader)
    previous_utility = [fed_valid_acc, fed_valid_loss]
    utility_dim = 2
    print("Previous utility: ", previous_utility)

    num_clients = len(client_models)
    shapley_value_all_rounds = [[] for _ in range(utility_dim)]
    shapley_value_sum = [{} for _ in range(utility_dim)]
    for i in range(utility_dim):
        shapley_value_all_rounds[i].append({cid: previous_utility[i] / num_clients for cid in range(num_clients)})
        shapley_value_sum[i] = shapley_value_all_rounds[i][0]
    print("shapley_value_all_rounds: {}".format(shapley_value_all_rounds))
    print("shapley_value_sum: {}".format(shapley_value_sum))

    if checkpoints is None and not getattr(opt, "synthetic", False):
        checkpoints = client_checkpoint_paths(num_clients)
    local_acc_all, local_loss_all = [], []
    client_model_all_rounds = [None] * num_clients
    client_model_selection_matrix = [False] * num_clients
    for i, client_model in enumerate(client_models):
        if checkpoints is not None:
            checkLocalTrainingModelExist(checkpoints[i])
            ckpt = torch.load(checkpoints[i], map_location="cpu")
            client_model.load_state_dict(ckpt["state_dict"])
            print("Model loaded!")
        accuracy, loss = evaluation(args, client_model, valid_loader)
        print("Accuracy: ", accuracy)
        print("Loss: ", loss)
        local_acc_all.append(accuracy)
        local_loss_all.append(loss)
        client_model_all_rounds[i] = get_difference_between_network_weights(client_model, init_global_model)
        client_model_selection_matrix[i] = True
    print("Local accuracy: ", local_acc_all)
    print("Local loss: ", local_loss_all)
    print("Client model selection matrix: ", client_model_selection_matrix)

    from shapley_vit_b200.synth import SizedStub

    train_sets = [dataset] * num_clients if client_sizes is None else [SizedStub(n) for n in client_sizes]
    clients_all = [ClientBase(cid, args, init_global_model, train_sets[cid]) for cid in range(num_clients)]
    server = ServerBase(args, init_global_model, clients_all, None, valid_loader, None)
    game = Game(clients_all, server, init_global_model, client_model_all_rounds, client_model_selection_matrix,
                previous_utility, utility_dim, args)
    LAST_SHAPLEY_VALUE = call_shapley_computation_method(args, game, None)
    return shapley_value_all_rounds, shapley_value_sum


def getOCTData2():
    """The reference loads a private OCT cell dataset through a module that is not in its
    repository (start.py:1, 51-56).  With --synthetic (or when that loader is unavailable) a
    synthetic dict-sample dataset of the same contract takes its place."""
    from shapley_vit_b200 import layout, synth

    if not (opt.synthetic or opt.synthetic_data):
        try:
            from .datasets.dataloader_cell import XrayDataLoader as CellDataLoader  # user-provided

            return CellDataLoader(root_dir=my_validation_dataset, mode="train", patch_size=opt.patch_size, sub_dir="")
        except ImportError as e:
            raise RuntimeError("shapleyserver/datasets/dataloader_cell.py is not part of the reference repository; "
                               "provide it, or run with --synthetic") from e
    cfg = layout.vit_preset(opt.vit_size, image=opt.image_size, n_cls=opt.num_classes)
    images, labels = synth.make_val_set(cfg, opt.val_size, opt.seed or 0)
    return synth.DictSampleDataset(images, labels)


def start():
    from shapley_vit_b200 import layout, synth

    opts.ensure_output_dir()
    dataset = getOCTData2()
    num_clients = opt.num_clients or opt.dist_num
    cfg = layout.vit_preset(opt.vit_size, image=opt.image_size, n_cls=opt.num_classes)
    def new_model():
        if opt.lora_rank > 0:   # the reference's get_peft_model(vit, LoraConfig(r=16, lora_alpha=8, ...)), start.py:274-276
            from shapley_vit_b200.models.vit import LoraViTForImageClassification
            return LoraViTForImageClassification(cfg, r=opt.lora_rank, lora_alpha=opt.lora_alpha, precision=opt.dtype)
        return ViTForImageClassification(cfg, precision=opt.dtype)

    init_global_model = new_model()
    client_sizes = None
    if opt.synthetic:
        seed = opt.seed or 0
        if opt.lora_rank > 0:
            w0, client_sds = synth.make_peft_state_dicts(cfg, num_clients, seed, r=opt.lora_rank, prefix="base_model.model.")
        else:
            w0 = synth.make_state_dict(cfg, seed)
            client_sds = [synth.make_client_state_dict(w0, j, seed) for j in range(num_clients)]
        init_global_model.load_state_dict(w0)
        client_models = []
        for sd in client_sds:
            m = new_model()
            m.load_state_dict(sd)
            client_models.append(m)
        client_sizes = synth.client_sizes(num_clients)
    else:
        if opt.loadModel:
            init_global_model.load_state_dict(torch.load(opt.loadModel, map_location="cpu")["state_dict"])
        client_models = [copy.deepcopy(init_global_model) for _ in range(num_clients)]
    print("Length of dataset: ", len(dataset))
    getInitialShapleyValue(dataset, init_global_model, *client_models, client_sizes=client_sizes)
    return LAST_SHAPLEY_VALUE


if __name__ == "__main__":
    start()
