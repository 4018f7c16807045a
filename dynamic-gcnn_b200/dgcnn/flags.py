"""dgcnn.flags -- configuration with the flag names, short names and defaults of
/root/reference/dgcnn/flags.py:6-167 (sub-commands train | inference | iotest).  Config only: no compute.

Differences, all deliberate: `--seed` is typed int and accepted by every sub-command;
`-io synthetic` selects an in-memory generator (h5py / larcv are not in this image); attributes keep the
reference's UPPER-CASE convention (flags.py:150-153).
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np


def _strtobool(v) -> int:
    s = str(v).strip().lower()
    if s in ("y", "yes", "t", "true", "on", "1"):
        return 1
    if s in ("n", "no", "f", "false", "off", "0"):
        return 0
    raise argparse.ArgumentTypeError("invalid truth value %r" % v)


class DGCNN_FLAGS:
    # model (flags.py:9-17)
    NUM_CLASS = 2
    MODEL_NAME = "dgcnn"
    TRAIN = True
    KVALUE = 20
    DEBUG = True
    EDGE_CONV_LAYERS = 3
    EDGE_CONV_FILTERS = 64
    FC_LAYERS = 2
    FC_FILTERS = "512,256"
    DTYPE = "f32"     # not in the reference: arithmetic of the 1x1 convolutions, f32 (fp32-faithful) | bf16
    # train / inference (flags.py:20-33)
    SEED = -1
    LEARNING_RATE = 0.001
    GPUS = [0]
    MINIBATCH_SIZE = 1
    WEIGHT_PREFIX = "./weights/snapshot"
    NUM_POINT = 2048
    NUM_CHANNEL = -1
    ITERATION = 10000
    REPORT_STEP = 100
    SUMMARY_STEP = 20
    CHECKPOINT_STEP = 500
    CHECKPOINT_NUM = 10
    CHECKPOINT_HOUR = 0.4
    # IO (flags.py:36-45)
    IO_TYPE = "h5"
    INPUT_FILE = "/scratch/kterao/dlprod_ppn_v08/dgcnn_p02_test_4ch.hdf5"
    OUTPUT_FILE = ""
    BATCH_SIZE = 1
    LOG_DIR = ""
    MODEL_PATH = ""
    DATA_KEY = "data"
    LABEL_KEY = "label"
    WEIGHT_KEY = ""
    SHUFFLE = 1

    # (short, long, type, class attribute holding the default, help)
    _COMMON = [
        ("-kv", "--kvalue", int, "KVALUE", "K value"),
        ("-db", "--debug", _strtobool, "DEBUG", "Extra verbose mode for debugging"),
        ("-ld", "--log_dir", str, "LOG_DIR", "Log dir"),
        ("-sh", "--shuffle", _strtobool, "SHUFFLE", "Shuffle the data entries"),
        ("-ecl", "--edge_conv_layers", int, "EDGE_CONV_LAYERS", "Number of edge-convolution layers"),
        ("-ecf", "--edge_conv_filters", str, "EDGE_CONV_FILTERS", "Number of filters in edge-convolution layers"),
        ("-fcl", "--fc_layers", int, "FC_LAYERS", "Number of fully-connected layers"),
        ("-fcf", "--fc_filters", str, "FC_FILTERS", "Number of filters in fully-connected layers"),
        ("-nc", "--num_class", int, "NUM_CLASS", "Number of classes"),
        ("-np", "--num_point", int, "NUM_POINT", "Point number"),
        ("-it", "--iteration", int, "ITERATION", "Iteration to run"),
        ("-bs", "--batch_size", int, "BATCH_SIZE", "Batch Size during training for updating weights"),
        ("-mbs", "--minibatch_size", int, "MINIBATCH_SIZE", "Mini-Batch Size during training for each GPU"),
        ("-rs", "--report_step", int, "REPORT_STEP", "Period (in steps) to print out loss and accuracy"),
        ("-mn", "--model_name", str, "MODEL_NAME", "model name identifier"),
        ("-mp", "--model_path", str, "MODEL_PATH", "model checkpoint file path"),
        ("-io", "--io_type", str, "IO_TYPE", "IO handler type"),
        ("-if", "--input_file", str, "INPUT_FILE", "comma-separated input file list"),
        ("-of", "--output_file", str, "OUTPUT_FILE", "output file name"),
        ("-dkey", "--data_key", str, "DATA_KEY", "A keyword to fetch data from file"),
        ("-lkey", "--label_key", str, "LABEL_KEY", "A keyword to fetch label from file"),
        ("-sd", "--seed", int, "SEED", "Seed for random number generators"),
        ("-dt", "--dtype", str, "DTYPE", "Arithmetic of the 1x1 convolutions: f32 (fp32-faithful) or bf16"),
    ]
    _TRAIN_ONLY = [
        ("-wp", "--weight_prefix", str, "WEIGHT_PREFIX", "Prefix (directory + file prefix) for snapshots of weights"),
        ("-lr", "--learning_rate", float, "LEARNING_RATE", "Initial learning rate"),
        ("-ss", "--summary_step", int, "SUMMARY_STEP", "Period (in steps) to store summary in tensorboard log"),
        ("-chks", "--checkpoint_step", int, "CHECKPOINT_STEP", "Period (in steps) to store snapshot of weights"),
        ("-chkn", "--checkpoint_num", int, "CHECKPOINT_NUM", "Number of the latest checkpoint to keep"),
        ("-chkh", "--checkpoint_hour", float, "CHECKPOINT_HOUR", "Period (in hours) to store checkpoint"),
    ]
    _WEIGHT_KEY = ("-wkey", "--weight_key", str, "WEIGHT_KEY", "A keyword to fetch weight from file")

    def __init__(self):
        self._build_parsers()

    def _add(self, parser, spec):
        short, long_, typ, attr, helpmsg = spec
        default = getattr(self, attr)
        if typ is str:
            default = str(default)
        parser.add_argument(short, long_, type=typ, default=default, help="%s [default: %s]" % (helpmsg, default))

    def _build_parsers(self):
        from .main_funcs import train, iotest, inference
        self.parser = argparse.ArgumentParser(description="Edge-GCNN Configuration Flags")
        sub = self.parser.add_subparsers(title="Modules", description="Valid subcommands", dest="script")
        self.train_parser = sub.add_parser("train", help="Train Edge-GCNN")
        self.inference_parser = sub.add_parser("inference", help="Run inference of Edge-GCNN")
        self.iotest_parser = sub.add_parser("iotest", help="Test iotools for Edge-GCNN")
        for spec in self._TRAIN_ONLY + [self._WEIGHT_KEY]:
            self._add(self.train_parser, spec)
        self._add(self.iotest_parser, self._WEIGHT_KEY)
        for p in (self.train_parser, self.inference_parser, self.iotest_parser):
            p.add_argument("--gpus", type=str, default="0", help="GPUs to utilize (comma-separated integers")
            for spec in self._COMMON:
                self._add(p, spec)
        self.train_parser.set_defaults(func=train)
        self.inference_parser.set_defaults(func=inference)
        self.iotest_parser.set_defaults(func=iotest)

    def parse_args(self, argv=None):
        args = self.parser.parse_args(argv)
        if not getattr(args, "func", None):
            self.parser.print_help()
            return None
        self.update(vars(args))
        print("\n\n-- CONFIG --")
        for name in sorted(vars(self)):
            if isinstance(getattr(self, name), argparse.ArgumentParser):
                continue
            print("%s = %r" % (name, getattr(self, name)))
        np.random.seed(self.SEED % (2 ** 32))
        import torch
        torch.manual_seed(self.SEED)
        return args.func(self)

    def update(self, args):
        for name, value in args.items():
            if name in ("func", "script"):
                continue
            setattr(self, name.upper(), value)
        if isinstance(self.GPUS, str):
            # the reference exports CUDA_VISIBLE_DEVICES here (flags.py:154); with one process per GPU the
            # launcher (torchrun) owns device assignment, so only export it when running single-process
            if "LOCAL_RANK" not in os.environ:
                os.environ.setdefault("CUDA_VISIBLE_DEVICES", self.GPUS)
            self.GPUS = [int(g) for g in self.GPUS.split(",")]
        if isinstance(self.INPUT_FILE, str):
            self.INPUT_FILE = [str(f) for f in self.INPUT_FILE.split(",")]
        for attr in ("EDGE_CONV_FILTERS", "FC_FILTERS"):
            v = getattr(self, attr)
            if isinstance(v, str):
                setattr(self, attr, [int(x) for x in v.split(",")] if v.find(",") > 0 else int(v))
        if int(self.SEED) < 0:
            self.SEED = int(time.time())
