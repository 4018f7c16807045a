"""dgcnn.main_funcs -- host driver loops behind `bin/dgcnn.py {train,inference,iotest}`:
/root/reference/dgcnn/main_funcs.py:16-305 re-expressed for one-process-per-GPU PyTorch.  Same batching
(batch -> micro-batch -> per-GPU slices, :139-152), same CSV schema (:108-111,184-191), same checkpoint naming
`WEIGHT_PREFIX-<iteration>` (:178-182) and resume-from-filename (:13-14,90-93).  Thin host code only.
"""
from __future__ import annotations

import datetime
import glob
import os
import sys
import time

import numpy as np
import torch


def round_decimals(val, digits):
    factor = float(np.power(10, digits))
    return int(val * factor + 0.5) / factor


def iteration_from_filename(file_name):
    return int((file_name.split("-"))[-1])


class Handlers:
    sess = None
    data_io = None
    csv_logger = None
    weight_io = None
    train_logger = None
    trainer = None
    iteration = 0


def iotest(flags):
    import dgcnn
    io = dgcnn.io_factory(flags)
    io.initialize()
    num_entries = io.num_entries()
    ctr = 0
    while ctr < num_entries:
        idx, data, label, weight = io.next()
        msg = str(ctr) + "/" + str(num_entries) + " ... " + str(idx) + " " + str(data[0].shape)
        if label is not None:
            msg += str(label[0].shape)
        if weight is not None:
            msg += str(weight[0].shape)
        print(msg)
        ctr += len(data)
    io.finalize()


def train(flags):
    flags.TRAIN = True
    handlers = prepare(flags)
    train_loop(flags, handlers)


def inference(flags):
    flags.TRAIN = False
    handlers = prepare(flags)
    inference_loop(flags, handlers)


def _maybe_init_distributed():
    import torch.distributed as dist
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend="nccl")


def prepare(flags):
    import dgcnn
    handlers = Handlers()
    if flags.BATCH_SIZE % (flags.MINIBATCH_SIZE * len(flags.GPUS)):
        msg = "--batch_size (%d) must be a modular of --gpus (%d) * --minibatch_size (%d)\n"
        sys.stderr.write(msg % (flags.BATCH_SIZE, len(flags.GPUS), flags.MINIBATCH_SIZE))
        sys.exit(1)
    _maybe_init_distributed()

    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not flags.TRAIN and getattr(flags, "OUTPUT_FILE", ""):
        # every rank would write the same file, each with only its own towers' predictions
        sys.stderr.write("inference with --output_file runs in ONE process (towers then run back to back on one GPU)\n")
        raise NotImplementedError

    handlers.data_io = dgcnn.io_factory(flags)
    handlers.data_io.initialize()
    handlers.data_io.next()

    flags.NUM_CHANNEL = handlers.data_io.num_channels()
    handlers.trainer = dgcnn.trainval(flags)
    handlers.trainer.initialize()
    handlers.weight_io = handlers.trainer          # .save / .restore stand in for tf.train.Saver
    if flags.TRAIN and getattr(flags, "WEIGHT_PREFIX", ""):
        save_dir = flags.WEIGHT_PREFIX[0:flags.WEIGHT_PREFIX.rfind("/")]
        if save_dir and not os.path.isdir(save_dir):
            os.makedirs(save_dir, exist_ok=True)

    handlers.iteration = 0
    loaded_iteration = 0
    if flags.MODEL_PATH:
        handlers.trainer.restore(flags.MODEL_PATH)
        loaded_iteration = iteration_from_filename(flags.MODEL_PATH)
        if flags.TRAIN:
            handlers.iteration = loaded_iteration + 1

    if flags.LOG_DIR and handlers.trainer._rank == 0:
        os.makedirs(flags.LOG_DIR, exist_ok=True)
        kind = "train" if flags.TRAIN else "inference"
        handlers.csv_logger = open("%s/%s_log-%07d.csv" % (flags.LOG_DIR, kind, loaded_iteration), "w")
    return handlers


def _slices(flags, arr, current_idx):
    """main_funcs.py:145-152: one MINIBATCH_SIZE slice per entry of flags.GPUS."""
    if arr is None:
        return None, current_idx
    out = []
    for _ in flags.GPUS:
        out.append(arr[current_idx:current_idx + flags.MINIBATCH_SIZE])
        current_idx += flags.MINIBATCH_SIZE
    return out, current_idx


def _prune_checkpoints(prefix, keep):
    import re
    pat = re.compile(re.escape(prefix) + r"-\d+$")           # only PREFIX-<iteration>: not prefix-100.bak, prefix-final
    files = sorted((f for f in glob.glob(prefix + "-*") if pat.match(f)), key=iteration_from_filename)
    for p in files[:-int(keep)] if keep and len(files) > int(keep) else []:
        os.remove(p)


def train_loop(flags, handlers):
    csv = handlers.csv_logger
    if csv:
        csv.write("iter,epoch,titer,ttrain,tio,tsave,tsummary,tsumiter,tsumtrain,tsumio,tsumsave,tsumsummary,"
                  "loss,accuracy,points_per_sec\n")
    tsum = tsum_train = tsum_io = tsum_save = tsum_summary = 0.0
    trainer = handlers.trainer
    while handlers.iteration < flags.ITERATION:
        tstamp = datetime.datetime.fromtimestamp(time.time()).strftime("%Y-%m-%d %H:%M:%S")
        tstart_iteration = time.time()
        it1 = handlers.iteration + 1
        report_step = flags.REPORT_STEP and (it1 % flags.REPORT_STEP == 0)
        checkpt_step = flags.CHECKPOINT_STEP and flags.WEIGHT_PREFIX and (it1 % flags.CHECKPOINT_STEP == 0)

        t0 = time.time()
        idx, data, label, weight = handlers.data_io.next()
        tspent_io = time.time() - t0
        tsum_io += tspent_io

        loss_v, accuracy_v = [], []
        trainer.zero_gradients(handlers.sess)
        t0 = time.time()
        current_idx = 0
        while current_idx < flags.BATCH_SIZE:                      # accumulate micro-batches
            data_v, nxt = _slices(flags, data, current_idx)
            label_v, _ = _slices(flags, label, current_idx)
            weight_v, _ = _slices(flags, weight, current_idx)
            current_idx = nxt
            res = trainer.accum_gradient(handlers.sess, data_v, label_v, weight_v, sync=False,
                                         last=current_idx >= flags.BATCH_SIZE)
            accuracy_v.append(res[1])
            loss_v.append(res[2])
        trainer.apply_gradient(handlers.sess)
        # loss / accuracy come back to the host once per iteration, after the optimizer step has been enqueued
        accuracy_v = [float(a) for a in accuracy_v]
        loss_v = [float(l) for l in loss_v]
        torch.cuda.synchronize()
        tspent_train = time.time() - t0
        tsum_train += tspent_train
        tspent_summary = 0.0

        if trainer._world > 1:
            # the tower mean the reference logs (trainval.py:59-60): loss / accuracy rode through the all-reduce
            loss = float(trainer.last_loss) / len(loss_v)
            accuracy = float(trainer.last_accuracy) / len(loss_v)
        else:
            loss, accuracy = float(np.mean(loss_v)), float(np.mean(accuracy_v))
        epoch = handlers.iteration * float(flags.BATCH_SIZE) / handlers.data_io.num_entries()
        tspent_save = 0.0
        if checkpt_step and trainer._rank == 0:
            t0 = time.time()
            path = trainer.save(flags.WEIGHT_PREFIX, handlers.iteration)
            _prune_checkpoints(flags.WEIGHT_PREFIX, flags.CHECKPOINT_NUM)
            tspent_save = time.time() - t0
            tsum_save += tspent_save
            print("saved @", path)
        tspent_iteration = time.time() - tstart_iteration
        tsum += tspent_iteration
        if csv:
            pps = flags.BATCH_SIZE * data.shape[1] / max(tspent_train, 1e-12)
            csv.write("%d,%g,%g,%g,%g,%g,%g,%g,%g,%g,%g,%g,%g,%g,%g\n" % (
                handlers.iteration, epoch, tspent_iteration, tspent_train, tspent_io, tspent_save, tspent_summary,
                tsum, tsum_train, tsum_io, tsum_save, tsum_summary, loss, accuracy, pps))
        if report_step and trainer._rank == 0:
            mem = torch.cuda.max_memory_allocated()
            print("Iteration %d (epoch %g) @ %s ... train time fraction %g%% max mem. %g ... loss %g accuracy %g" % (
                handlers.iteration, round_decimals(epoch, 2), tstamp,
                round_decimals(tspent_train / tspent_iteration * 100.0, 2), mem, round_decimals(loss, 4),
                round_decimals(accuracy, 4)))
            sys.stdout.flush()
            if csv:
                csv.flush()
        handlers.iteration += 1
    if csv:
        csv.close()
    handlers.data_io.finalize()


def inference_loop(flags, handlers):
    csv = handlers.csv_logger
    if csv:
        csv.write("iter,epoch,titer,tinference,tio,tsumiter,tsuminference,tsumio,loss,accuracy\n")
    tsum = tsum_io = tsum_inference = 0.0
    trainer = handlers.trainer
    while handlers.iteration < flags.ITERATION:
        tstamp = datetime.datetime.fromtimestamp(time.time()).strftime("%Y-%m-%d %H:%M:%S")
        tstart_iteration = time.time()
        report_step = flags.REPORT_STEP and ((handlers.iteration + 1) % flags.REPORT_STEP == 0)
        t0 = time.time()
        idx, data, label, weight = handlers.data_io.next()
        tspent_io = time.time() - t0
        tsum_io += tspent_io

        softmax_vv, loss_v, accuracy_v = [], [], []
        t0 = time.time()
        current_idx = 0
        while current_idx < flags.BATCH_SIZE:
            data_v, nxt = _slices(flags, data, current_idx)
            label_v, _ = _slices(flags, label, current_idx)
            weight_v, _ = _slices(flags, weight, current_idx)
            current_idx = nxt
            res = trainer.inference(handlers.sess, data_v, label_v, weight_v)
            if label_v is not None:
                softmax_vv += res[0:-2]
                accuracy_v.append(res[-2])
                loss_v.append(res[-1])
            else:
                softmax_vv += res
        tspent_inference = time.time() - t0
        tsum_inference += tspent_inference

        if flags.OUTPUT_FILE:
            # softmax_vv holds, per micro-step, one array per tower this process ran (all of them: single process)
            idx_ctr = 0
            for softmax_v in softmax_vv:
                for softmax in softmax_v:
                    handlers.data_io.store(idx[idx_ctr], softmax)
                    idx_ctr += 1
        loss, accuracy = (-1, -1)
        if loss_v:
            loss, accuracy = float(np.mean(loss_v)), float(np.mean(accuracy_v))
        epoch = handlers.iteration * float(flags.BATCH_SIZE) / handlers.data_io.num_entries()
        tspent_iteration = time.time() - tstart_iteration
        tsum += tspent_iteration
        if csv:
            csv.write("%d,%g,%g,%g,%g,%g,%g,%g,%g,%g\n" % (handlers.iteration, epoch, tspent_iteration,
                                                         tspent_inference, tspent_io, tsum, tsum_inference, tsum_io,
                                                         loss, accuracy))
        if report_step:
            print("Iteration %d (epoch %g) @ %s ... inference time fraction %g%% max mem. %g ... loss %g accuracy %g"
                  % (handlers.iteration, round_decimals(epoch, 2), tstamp,
                     round_decimals(tspent_inference / tspent_iteration * 100.0, 2), torch.cuda.max_memory_allocated(),
                     round_decimals(loss, 4), round_decimals(accuracy, 4)))
            sys.stdout.flush()
            if csv:
                csv.flush()
        handlers.iteration += 1
    if csv:
        csv.close()
    handlers.data_io.finalize()
