"""Entry points of the reference ``train_clip_final.py`` for the EEG-CLIP hot path, on B200 kernels.

Call-compatible pieces:
  * ``load_eeg_encoder`` / ``load_speech_encoder``  -- same positional signatures and hard-wired hyper-parameters
    as train_clip_final.py:37-100 / 102-130 for the encoders in scope;
  * the CLI flags of train_clip_final.py:158-216 (same names and defaults), parsed by ``build_parser``;
  * ``train_step`` -- the body of the hot loop (train_clip_final.py:476-492): forward, zero_grad, backward, AdamW;
  * ``main`` -- the training loop with StepLR per epoch (:419,503-504), validation and checkpointing (:506-540),
    optionally data-parallel under torchrun (sharded InfoNCE + SUM all-reduce; new functionality).

Dataset plumbing (dataset_loader.py, the file split) is out of scope: ``main`` is wired for
``--data_dir synthetic`` only (seeded synthetic batches of the reference's shapes) and raises for any other value.
A caller with the reference's ``EEGDatasetSimdata`` at hand drives ``build_model`` / ``train_step`` / ``DevicePrefetcher``
with its own loader -- they take the reference's batch tuples as they are.
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist

from . import _lib as L
from .clip_model import (CLIPKLDNoLatentProj, CLIPNoContrastiveLearning, CLIPSim, CLIPSimMultiplePositives, CLIPSimNoLatentProj,
                         EEGConformer, EEGConformerInterleaved, EEGConvLSTM, SpeechSmallConv, memoryBank)
from .optim import Adam, AdamW
from .parallel import BucketedGradReducer, allreduce_gradients, bind_to_gpu_numa_node, broadcast_parameters
from .vlaai import VLAAI

# flag table: (name, type, default, choices) -- train_clip_final.py:163-216
_FLAGS = [
    ("debug", str, "no", ["yes", "no"]), ("only_evaluate", str, "no", ["yes", "no"]),
    ("results_folder", str, os.path.join(os.path.dirname(os.path.abspath(__file__)), "results"), None),
    ("run", int, 4, None), ("lstm_units", int, 128, None), ("lambda_sim_loss", float, 0.0, None),
    ("warmup_epochs", int, 0, None), ("momentum_membank", float, 0.90, None), ("eeg_norm", str, "mvn", ["mvn"]),
    ("stimulus_features", str, "wav2vec_19", None),
    ("model_arch", str, "clip_sim_no_latent_proj",
     ["no_contrastive_learning", "clip_kld", "clip_kld_latent_proj", "clip_mp", "clip_sim", "clip_sim_no_latent_proj",
      "clip_extended", "clip_no_eeg_loss", "clip_correct"]),
    ("speech_encoder", str, "convLSTM", ["conformer", "smallConv", "lstm", "convLSTM", "no", "double_lstm", "Wav2vecSmallModel"]),
    ("eeg_encoder", str, "EEGConformerInterleaved",
     ["EEGConformerInterleaved", "conformer", "convLSTMnew", "convLSTM", "lstm_newvals", "vlaai", "clipmeta", "lstm", "lstm_lstm",
      "double_lstm", "transformerEncoder"]),
    ("attention_depth", int, 10, None), ("load_pretrain", str, "no", ["yes", "no"]), ("shuffle", str, "yes", ["yes", "no"]),
    ("shuffle_percentage", float, 1.0, None), ("addEEG", str, "no", ["yes", "no"]),
    ("data_augmentation", str, "no", ["no", "SignFlip", "FTSurrogate", "FrequencyShift", "BandstopFilter", "GaussianNoise",
                                      "SmoothTimeMask", "ChannelsDropout", "ChannelsShuffle"]),
    ("data_augmentation_percentage", float, 0.5, None), ("learning_rate", float, 1e-3, None), ("beta1", float, 0.90, None),
    ("beta2", float, 0.999, None), ("use_amsgrad", str, "no", ["yes", "no"]), ("optimizer", str, "adamw", ["adam", "adamw"]),
    ("weight_decay", float, 0.01, None), ("lr_scheduler", str, "step", ["no", "plateau", "step", "cosine", "cosine_warmup"]),
    ("step_size_scheduler", int, 10, None), ("epochs", int, 500, None), ("patience", int, 15, None),
    ("batch_size", int, 128, None), ("number_conv_layers", int, 1, None), ("fun_act", str, "relu", None),
    ("temperature", float, 0.075, None), ("subject_split", str, "icassp_testset", ["within", "heldout", "icassp_testset"]),
    ("data_dir", str, "/esat/audioslave/lbollens/sparrkulee_data/sparrkulee", None),
    ("number_of_training_subjects", int, 1000, None), ("lambda_clip_loss", float, 1, None), ("latent_dim", int, 8, None),
]

WINDOW_LENGTH = 3 * 64  # train_clip_final.py:150-152


def build_parser():
    p = argparse.ArgumentParser(description="Train CLIP model (B200 kernels).")
    for name, typ, default, choices in _FLAGS:
        p.add_argument(f"--{name}", type=typ, default=default, choices=choices)
    # additions (not in the reference): synthetic-data sizing for --data_dir synthetic
    p.add_argument("--window_length", type=int, default=WINDOW_LENGTH)
    p.add_argument("--synthetic_batches", type=int, default=8)
    p.add_argument("--math", type=str, default=None, choices=["fp32", "bf16x3", "bf16"])
    return p


_latent_dim_global = 8  # the reference's smallConv / convLSTM speech factories read the *global* latent_dim (:114,119)


def load_eeg_encoder(eeg_encoder, units_lstm, padding, spatial_filters, number_conv_layers, window_length, latent_dim,
                     attention_depth):
    """train_clip_final.py:37-100 for the encoders on the hot path."""
    conv = dict(dropout_rate=0.2, eeg_dim=64, filters=(64,) * number_conv_layers, kernels=(64,) * number_conv_layers,
                dilation_rate=1, input_channels=64, time_dimension=window_length)
    if eeg_encoder == 'EEGConformerInterleaved':
        return EEGConformerInterleaved(output_dim=latent_dim, conformer_input_dim=64, depth=attention_depth, **conv)
    if eeg_encoder == 'conformer':
        return EEGConformer(output_dim=latent_dim, conformer_input_dim=64, depth=attention_depth, **conv)
    if eeg_encoder == 'vlaai':
        return VLAAI()
    if eeg_encoder == 'convLSTM':
        return EEGConvLSTM(units_lstm=128, output_dim=latent_dim, dropout_rate=0.4, eeg_dim=64, filters=(64,) * number_conv_layers,
                           kernels=(32,) * number_conv_layers, dilation_rate=1, input_channels=64, time_dimension=window_length,
                           normalization_fn='layer_norm', activation_fn='leaky_relu')
    # same outcome as the reference for names it never constructs: `eeg` is unbound (:100)
    raise UnboundLocalError(f"eeg encoder '{eeg_encoder}' is outside the B200 hot path (SURVEY §2.1 #9)")


def load_speech_encoder(speech_encoder, units_lstm, padding, spatial_filters, number_conv_layers, window_length,
                        stride_temporal, speech_dimension):
    """train_clip_final.py:102-130 for the encoders on the hot path."""
    if speech_encoder == 'smallConv':
        return SpeechSmallConv(output_dim=_latent_dim_global, ks_temporal=16, dropout_rate=0.4, speech_dim=speech_dimension,
                               time_dimension=window_length)
    if speech_encoder == 'convLSTM':
        return EEGConvLSTM(units_lstm=128, output_dim=_latent_dim_global, dropout_rate=0.4, eeg_dim=speech_dimension,
                           filters=(64,) * number_conv_layers, kernels=(32,) * number_conv_layers, dilation_rate=1,
                           input_channels=speech_dimension, time_dimension=window_length, normalization_fn='layer_norm',
                           activation_fn='leaky_relu')
    raise UnboundLocalError(f"speech encoder '{speech_encoder}' is outside the B200 hot path (SURVEY §2.1 #8)")


def speech_dimension_of(stimulus_features):
    """train_clip_final.py:292-300."""
    if stimulus_features == 'mel':
        return 28, 64
    if stimulus_features == 'env':
        return 1, 8
    if 'wav2vec' in stimulus_features:
        return 1024, 128
    raise ValueError(stimulus_features)


def build_model(args, window_length, bank_size, device):
    """Model construction of train_clip_final.py:338-399."""
    global _latent_dim_global
    _latent_dim_global = args.latent_dim
    speech_dim, spatial_filters = speech_dimension_of(args.stimulus_features)
    eeg = load_eeg_encoder(args.eeg_encoder, args.lstm_units, 'valid', spatial_filters, args.number_conv_layers, window_length,
                           args.latent_dim, args.attention_depth)
    speech = load_speech_encoder(args.speech_encoder, args.lstm_units, 'valid', spatial_filters, args.number_conv_layers,
                                 window_length, 3, speech_dim)
    arch = args.model_arch
    latent_dim = args.latent_dim
    if arch in ('clip_sim_no_latent_proj', 'clip_kld'):
        latent_dim = speech.get_output_dim(window_length)      # bank rows / KLD input are the flattened tower output (:367-368)
    bank = memoryBank(bank_size=bank_size, dim=latent_dim, momentum=args.momentum_membank, device=device) if bank_size else None
    common = dict(temperature=args.temperature, window_length=window_length, lambda_clip=args.lambda_clip_loss)
    if arch == 'clip_sim_no_latent_proj':
        model = CLIPSimNoLatentProj(eeg, speech, bank, lambda_average=args.lambda_sim_loss, **common)
    elif arch == 'clip_sim':
        model = CLIPSim(eeg, speech, bank, latent_dim=latent_dim, lambda_average=args.lambda_sim_loss, **common)
    elif arch == 'clip_mp':
        model = CLIPSimMultiplePositives(eeg, speech, lambda_average=args.lambda_sim_loss, **common)
    elif arch == 'clip_kld':
        model = CLIPKLDNoLatentProj(eeg, speech, latent_dimension=latent_dim, number_of_classes=bank_size,
                                    lambda_lower_bound=args.lambda_sim_loss, lambda_discriminative=args.lambda_sim_loss, **common)
    elif arch == 'no_contrastive_learning':
        model = CLIPNoContrastiveLearning(eeg, speech, window_length=window_length)
    else:
        # same outcome as the reference for the names its factory never constructs: `model` is unbound (:379-399)
        raise NameError(f"model_arch '{arch}' is not constructed by train_clip_final.py:379-396")
    return model.to(device)


def build_optimizer(args, params):
    """train_clip_final.py:403-413."""
    kw = dict(betas=(args.beta1, args.beta2), amsgrad=args.use_amsgrad == 'yes', lr=args.learning_rate)
    if args.optimizer == 'adam':
        return Adam(params, **kw)
    return AdamW(params, weight_decay=args.weight_decay, **kw)


def train_step(model, optimizer, eeg, speech, ids, use_total=True, group=None):
    """One iteration of the reference hot loop (train_clip_final.py:484-492).

    Order is the reference's: forward, zero_grad, backward, step.  Under a data-parallel ``group`` gradients are
    SUM-all-reduced (one collective on the optimizer's flat arena) before the step.
    """
    out = model(eeg, speech, ids)
    if len(out) == 4:                                  # the KLD wrappers return (loss_total, loss_ce, log_pmu2, kld_z2) (:1277)
        loss_total, loss_ce, loss_avg = out[0], out[1], out[3]
    else:
        loss_ce, loss_avg, loss_total = out
    optimizer.zero_grad()
    if group is not None and dist.is_initialized() and dist.get_world_size(group) > 1 and hasattr(optimizer, "flat_grads"):
        # bucketed: the EEG tower's gradients go out while the speech tower's backward runs, the rest after the backward
        with BucketedGradReducer(optimizer, group) as reducer:
            (loss_total if use_total else loss_ce).backward()
            reducer.finish()
    else:
        (loss_total if use_total else loss_ce).backward()
        if group is not None:
            allreduce_gradients(model.parameters(), group)
    optimizer.step()
    return loss_ce, loss_avg, loss_total


class SyntheticBatches:
    """Seeded stand-in for EEGDatasetSimdata (dataset_loader.py:392-422): yields (eeg, [speech], ids, subs)."""

    def __init__(self, n_batches, batch_size, window_length, speech_dim, seed=0, n_segments=10000):
        self.n, self.b, self.T, self.F, self.seed, self.n_segments = n_batches, batch_size, window_length, speech_dim, seed, n_segments

    def get_number_of_stimuli_segments(self):
        return self.n_segments

    def __iter__(self):
        g = torch.Generator().manual_seed(self.seed)
        for _ in range(self.n):
            eeg = torch.randn(self.b, self.T, 64, generator=g)
            speech = torch.randn(self.b, self.T, self.F, generator=g)
            ids = torch.randperm(self.n_segments, generator=g)[:self.b] + 1
            yield eeg, [speech], ids, torch.zeros(self.b, dtype=torch.int64)


class DevicePrefetcher:
    """Input pipeline of the hot loop (train_clip_final.py:475-479): yields device-resident (eeg, speech, ids) while the
    NEXT batch's host->device copies run on a side stream, so the 201-357 MB of wav2vec2 features per batch cross PCIe
    under the previous step instead of in front of it.  Two persistent device buffers (and two pinned staging buffers when
    the loader's tensors are pageable) are reused in turn: no allocation in steady state, ordering by CUDA events only.
    `loader` yields the reference's batch tuples (eeg (b,T,64), [speech (b,T,F)], ids (b,), subs).  A yielded batch is
    valid until the next-but-one batch is requested."""

    def __init__(self, loader, device, speech_stage_dtype=None):
        """``speech_stage_dtype=torch.float16`` halves the host->device traffic of the wav2vec2 features (the step's largest
        transfer: 335 MB per 256 windows; at 8 ranks per host the fp32 stream is pinned-memory bound).  The features are rounded
        to fp16 on the host (2.4e-4 relative) and widened back to fp32 on the device, so this is an opt-in deviation from the
        reference's fp32 inputs; the default (None) moves fp32."""
        self.loader, self.device = loader, device
        self.speech_stage_dtype = speech_stage_dtype
        self.stream = torch.cuda.Stream(device=device)
        self._pinned = [None, None]
        self._dev = [None, None]
        self._wide = [None, None]
        self._free = [None, None]      # main-stream event after which device buffer `slot` may be overwritten
        self._copied = [None, None]    # copy-stream event after which pinned staging buffer `slot` may be rewritten by the host

    def _stage(self, slot, data):
        eeg, speech, ids = data[0], data[1][0] if isinstance(data[1], (list, tuple)) else data[1], data[2]
        sp_dtype = self.speech_stage_dtype or torch.float32
        src = (eeg.float(), speech.to(sp_dtype), ids.to(torch.int64))
        if all(t.is_pinned() for t in src):
            pin = src                                  # the loader already hands out page-locked batches (DataLoader(pin_memory=True))
        else:
            pin = self._pinned[slot]
            if pin is None or any(p.shape != t.shape for p, t in zip(pin, src)):
                pin = self._pinned[slot] = tuple(torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in src)
            if self._copied[slot] is not None:
                self._copied[slot].synchronize()       # the previous H2D copy OUT of this pinned buffer has finished (host-side wait:
                                                       # stream events order the device, not the host's next write into the buffer)
            for p_, t in zip(pin, src):
                p_.copy_(t)
        dev = self._dev[slot]
        if dev is None or any(d.shape != t.shape or d.dtype != t.dtype for d, t in zip(dev, src)):
            dev = self._dev[slot] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in src)
            self._wide[slot] = None if sp_dtype == torch.float32 else torch.empty(src[1].shape, dtype=torch.float32, device=self.device)
        with torch.cuda.stream(self.stream):
            if self._free[slot] is not None:
                self.stream.wait_event(self._free[slot])   # the step that read this buffer has finished
            for d, p_ in zip(dev, pin):
                d.copy_(p_, non_blocking=True)
            out = dev
            if self._wide[slot] is not None:           # widen on the copy stream into a persistent fp32 buffer
                self._wide[slot].copy_(dev[1])
                out = (dev[0], self._wide[slot], dev[2])
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._copied[slot] = ev
        return out, ev

    def __iter__(self):
        it = iter(self.loader)
        main = torch.cuda.current_stream(self.device)
        slot = 0
        try:
            nxt = self._stage(slot, next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            cur_slot = slot
            slot ^= 1
            try:
                nxt = self._stage(slot, next(it))      # overlaps the step that consumes `cur`
            except StopIteration:
                nxt = None
            main.wait_event(ev)
            yield cur
            done = torch.cuda.Event()                  # the consumer's work on `cur` is enqueued by now
            done.record(main)
            self._free[cur_slot] = done


class LossReader:
    """Device -> host read of every step's loss without draining the launch queue: the 0-dim loss is copied into a pinned
    ring slot on the compute stream (non-blocking) and read on the host ONE step later, when its event has long fired.
    ``loss.item()`` right after a step (the reference does it every 100 batches, train_clip_final.py:494-500) makes the host
    wait for the whole step and the GPU then idles while the next step's ~500 launches are enqueued."""

    def __init__(self, device, depth=4):
        self.buf = torch.empty(depth, dtype=torch.float32).pin_memory()
        self.ev = [torch.cuda.Event() for _ in range(depth)]
        self.depth, self.n_push, self.n_pop, self.device = depth, 0, 0, device

    def push(self, loss):
        """Enqueue the copy of this step's loss; returns the losses of earlier steps that are due (lag one step)."""
        out = []
        while self.n_push > self.n_pop:              # the previous step's event: this step's launches are already queued
            out.append(self._pop())
        slot = self.n_push % self.depth
        self.buf[slot:slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        self.ev[slot].record(torch.cuda.current_stream(self.device))
        self.n_push += 1
        return out

    def _pop(self):
        slot = self.n_pop % self.depth
        self.ev[slot].synchronize()
        self.n_pop += 1
        return float(self.buf[slot])

    def drain(self):
        return [self._pop() for _ in range(self.n_push - self.n_pop)]


def printf(s, file):
    print(s)
    with open(file, 'a') as f:
        f.write(s + '\n')


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.math:
        L.set_default_math(args.math)
    if not torch.cuda.is_available():
        raise L.EegclipError("train_clip_final (B200): no CUDA device; this implementation has no CPU path")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    window_length = args.window_length
    speech_dim, _ = speech_dimension_of(args.stimulus_features)
    if args.data_dir != "synthetic":
        raise L.EegclipError("only --data_dir synthetic is wired here; dataset plumbing is the reference's own (out of scope)")
    train_data = SyntheticBatches(args.synthetic_batches, args.batch_size, window_length, speech_dim, seed=rank)
    val_data = SyntheticBatches(max(1, args.synthetic_batches // 4), args.batch_size, window_length, speech_dim, seed=1000 + rank)
    model = build_model(args, window_length, train_data.get_number_of_stimuli_segments(), device)
    model.shard_group = group
    broadcast_parameters(model, group)
    opt = build_optimizer(args, model.parameters())
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=args.step_size_scheduler, gamma=0.1) if args.lr_scheduler == 'step' else None
    results = os.path.join(args.results_folder, f"results_{args.model_arch}_eeg_{args.eeg_encoder}_speech_{args.speech_encoder}_date_"
                           f"{time.strftime('%m-%d-%H-%M-%S')}")
    ckpt_dir = os.path.join(results, 'checkpoints')
    file_loss = os.path.join(results, 'loss.txt')
    if rank == 0:
        os.makedirs(ckpt_dir, exist_ok=True)
        with open(os.path.join(results, 'args.txt'), 'w') as f:
            json.dump(args.__dict__, f, indent=2)
    best_loss, best_epoch = float('inf'), 0
    for epoch in range(args.epochs):
        if epoch > best_epoch + args.patience and epoch > args.warmup_epochs:
            break
        model.train()
        for batch, (eeg, speech, ids) in enumerate(DevicePrefetcher(train_data, device)):
            loss_ce, loss_avg, _ = train_step(model, opt, eeg, speech, ids, use_total=epoch >= args.warmup_epochs, group=group)
            if batch % 100 == 0 and rank == 0:
                printf(f'train epoch {epoch} batch {batch} loss_ce  {loss_ce.item()} loss average eeg {loss_avg.item()}', file_loss)
        if sched is not None:
            sched.step()
        model.eval()
        ce = []
        with torch.no_grad():
            for data in val_data:
                l_ce, _, _ = model(data[0].to(device), data[1][0].to(device), data[2].to(device, dtype=torch.int64))
                ce.append(l_ce)
        mean_ce = torch.stack(ce).mean().item()
        if rank == 0:
            printf(f'validation epoch {epoch}: mean loss ce : {mean_ce}', file_loss)
            if mean_ce < best_loss:
                torch.save(model.state_dict(), os.path.join(ckpt_dir, 'model.ckpt'))
        if mean_ce < best_loss:
            best_loss, best_epoch = mean_ce, epoch
    if world > 1:
        dist.destroy_process_group()
    return best_loss


if __name__ == '__main__':
    main()
