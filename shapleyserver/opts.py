"""Command-line options: every flag of the reference's shapleyserver/opts.py (:16-90) with the
same names, aliases, types and defaults, plus additive flags for the batched utility loop.

As in the reference the flags are parsed when the module is imported and exposed as
``opts`` / ``opt``.  Differences: unknown arguments are tolerated (so the module can be imported
under a test runner), and ``exp/<exp_id>/`` is created by ``Opts.ensure_output_dir()`` (called
from ``start()``) instead of as an import side effect.
"""
from __future__ import annotations

import argparse
import os
import sys

_STORE_TRUE = "store_true"

# (flags, kwargs) -- the reference's option table
_REFERENCE_FLAGS = [
    (("--dist-num", "--dist_num"), dict(type=int, default=3, help="number of parties")),
    (("--dist-rank", "--dist_rank"), dict(type=int, default=0, help="rank of parties")),
    (("--master-addr", "--master_addr"), dict(type=str, default="172.20.117.210", help="master address")),
    (("--master-port", "--master_port"), dict(type=int, default=29500, help="master port")),
    (("--exp-id", "--exp_id"), dict(default="default", help="Experiment ID")),
    (("--exp-dir", "--exp_dir"), dict(default="exp", help="Experiment directory")),
    (("-test",), dict(action=_STORE_TRUE, help="test")),
    (("-demo",), dict(default="", help="path/to/demo/image")),
    (("-resume",), dict(default=False, type=bool, metavar="BOOL", help="Use the checkpoint or not")),
    (("-loadModel",), dict(default=None, help="Provide full path to a previously trained model")),
    (("-lr",), dict(type=float, default=3e-1, help="Learning Rate")),
    (("-epochs",), dict(type=int, default=250, help="#training epochs")),
    (("-trainBatch",), dict(type=int, default=8, help="Mini-batch size")),
    (("--batch-size", "--batch_size"), dict(type=int, default=32, help="batch size")),
    (("--clear-cache", "--clear_cache"), dict(default=False, type=bool, metavar="BOOL", help="Clear dataset cache")),
    (("--plot-server", "--plot_server"), dict(type=str, default="http://10.10.10.100", help="IP address")),
    (("--exp-name", "--exp_name"), dict(type=str, default="lstm_gaze", help="The env name in visdom")),
    (("--plot-port", "--plot_port"), dict(type=int, default=31831, help="Port number")),
    (("--save-interval", "--save_interval"), dict(type=int, default=1, help="save interval")),
    (("--snapshot-fname-prefix", "--snapshot_fname_prefix"), dict(default="", type=str, metavar="PATH", help="path to snapshot")),
    (("--sal-image-fname-dir", "--sal_image_fname_dir"), dict(default="exps/", type=str, metavar="PATH", help="path to sal image")),
    (("--epoch-st", "--epoch_st"), dict(default=0, type=int, help="first epoch")),
    (("--epoch-end", "--epoch_end"), dict(default=250, type=int, help="last epoch")),
    (("--debug",), dict(dest="debug", action=_STORE_TRUE, default=False, help="debug")),
    (("--eval",), dict(dest="eval", action=_STORE_TRUE, default=False, help="evaluate only")),
    (("--use-vis", "--use_vis"), dict(dest="use_vis", action=_STORE_TRUE, default=False, help="use vis")),
    (("--mode",), dict(type=str, default="train", help="mode name")),
    (("--patch-size", "--patch_size"), dict(type=int, default=256, help="patch size for train")),
    (("--data-dir", "--data_dir"), dict(type=str, default="/media/astar/e006bf52-80e3-47f3-b5ca-f5871c5a5e7f/home/astar/FL_Platform/OCT/CellData/OCT/", help="dataset directory")),
    (("--data-sub-dir", "--data_sub_dir"), dict(type=str, default=None, help="dataset sub dir")),
    (("--model-type", "--model_type"), dict(type=str, default="ViT", help="model type")),
    (("--use-grad-cam", "--use_grad_cam"), dict(dest="use_grad_cam", action=_STORE_TRUE, default=False, help="use grad cam")),
    (("--use-tensorboard", "--use_tensorboard"), dict(dest="use_tensorboard", action=_STORE_TRUE, default=False, help="use tensorboard")),
    (("--use-grad-cam-layers", "--use_grad_cam_layers"), dict(dest="use_grad_cam_layers", action=_STORE_TRUE, default=False, help="use grad cam layers")),
    (("--epsilon",), dict(type=float, default=0, help="epsilon")),
    (("--adv-dataset-mode", "--adv_dataset_mode"), dict(type=str, default="train", help="adv dataset mode")),
    (("--requires-control", "--requires_control"), dict(dest="requires_control", action=_STORE_TRUE, default=False, help="requires control")),
    (("--is-defense", "--is_defense"), dict(dest="is_defense", action=_STORE_TRUE, default=False, help="is defense")),
    (("--use-clean-eval", "--use_clean_eval"), dict(dest="use_clean_eval", action=_STORE_TRUE, default=False, help="use clean eval")),
    (("--use-multi-epsilon", "--use_multi_epsilon"), dict(dest="use_multi_epsilon", action=_STORE_TRUE, default=False, help="use multi epsilon")),
    (("--dataset-type", "--dataset_type"), dict(type=str, default="x-ray", help="dataset type")),
    (("--num-of-tasks", "--num_of_tasks"), dict(type=int, default=14, help="number of tasks")),
    (("--use-whole-dataset", "--use_whole_dataset"), dict(dest="use_whole_dataset", action=_STORE_TRUE, default=False, help="use whole dataset")),
    (("--noise-multiplier", "--noise_multiplier"), dict(type=float, default=0.5, help="dp noise multiplier")),
]

# additive flags (not in the reference) that reach the utility loop
_NEW_FLAGS = [
    (("--approximation-method", "--approximation_method"), dict(type=str, default="comp_contrib",
        help="comp_contrib (reference default) | exact | exact_own | monte_carlo | gtg | mr | tmr | group_testing")),
    (("--num-clients", "--num_clients"), dict(type=int, default=None, help="number of clients (default: --dist-num)")),
    (("--seed",), dict(type=int, default=None, help="seed for the estimators' RNG streams")),
    (("--mc-samples", "--mc_samples"), dict(type=int, default=None, help="samples m (default 50 n / 100)")),
    (("--coalition-batch", "--coalition_batch"), dict(type=int, default=8, help="coalitions per grouped-GEMM batch")),
    (("--image-chunk", "--image_chunk"), dict(type=int, default=128, help="validation images per forward")),
    (("--vit-size", "--vit_size"), dict(type=str, default="base", help="tiny | small | base | large")),
    (("--image-size", "--image_size"), dict(type=int, default=224, help="ViT input resolution")),
    (("--num-classes", "--num_classes"), dict(type=int, default=4, help="classifier width (reference: 4)")),
    (("--dtype",), dict(type=str, default="f16c8", help="GEMM arithmetic: f16c8 (default: fp16 + fp8-compensated, meets the 99.9 % top-1 parity gate) | f16x3 | f32 (parity modes) | f16 | bf16 | tf32 (throughput modes, outside the gate)")),
    (("--lora-rank", "--lora_rank"), dict(type=int, default=0,
                                          help="> 0: PEFT-LoRA wrapped ViT (query / value, classifier saved) as in the "
                                               "reference's start.py:274-276 (there: 16); 0: plain ViT")),
    (("--lora-alpha", "--lora_alpha"), dict(type=float, default=8.0, help="LoRA alpha (reference: 8)")),
    (("--synthetic",), dict(action=_STORE_TRUE, default=False, help="synthetic validation set and client models")),
    (("--synthetic-data", "--synthetic_data"), dict(action=_STORE_TRUE, default=False,
        help="synthetic validation set only: the client models come from the reference's checkpoint layout "
             "(shapleyserver/local_training/client_<i>_model/ViT_epoch_9.pth.tar under the working directory), the initial "
             "global model from -loadModel")),
    (("--val-size", "--val_size"), dict(type=int, default=1000, help="synthetic validation images")),
]


class Opts():
    def __init__(self, argv=None):
        self.parser = argparse.ArgumentParser()
        self.init()
        self.opt, self.unknown = self.parser.parse_known_args(sys.argv[1:] if argv is None else argv)
        self.opt.output_dir = os.path.join(self.opt.exp_dir, self.opt.exp_id)

    def init(self):
        self.parser.add_argument("--fl", dest="no_fl", action="store_false", help="use fl")
        self.parser.add_argument("--no-fl", dest="no_fl", action="store_true", help="no fl")
        self.parser.set_defaults(no_fl=True)
        for flags, kw in _REFERENCE_FLAGS + _NEW_FLAGS:
            self.parser.add_argument(*flags, **kw)

    def ensure_output_dir(self):
        os.makedirs(self.opt.output_dir, exist_ok=True)
        return self.opt.output_dir

    def log(self, printer=print):
        printer("\nArgs:")
        for k, v in sorted(vars(self.opt).items()):
            printer("%s,%s" % (str(k), str(v)))


opts = Opts()
opt = opts.opt
