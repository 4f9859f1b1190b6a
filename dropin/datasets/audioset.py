"""Synthetic stand-in for the reference's `datasets/audioset.py`.

The reference module refuses to import until `dataset_dir` points at three multi-hundred-GB HDF5 files of mp3
bytes (datasets/audioset.py:19-22) and decodes them with PyAV + h5py (neither is installed, and there is no
network).  This module offers the same three entry points `ex_audioset.py` uses --

    get_full_training_set(add_index, roll, wavmix, gain_augment, resample_rate)   datasets/audioset.py:255-264
    get_test_set(resample_rate)                                                   datasets/audioset.py:267-269
    get_ft_weighted_sampler(epoch_len, sampler_replace)                           datasets/audioset.py:175-177

-- over deterministic synthetic clips with the same item layout:
    train item  (waveform [1, N] float32, audio_name str, target [527] float32, index int)   (AddIndexDataset, :99-108)
    test  item  (waveform [1, N] float32, audio_name str, target [527] float32)
Clip i is N(0, 0.1^2) noise from a per-clip seeded generator plus two label-dependent tones, so that it neither
depends on the batch size nor on the order of access.  Every class has at least one positive and one negative
clip in the test split (sklearn's roc_auc_score needs both).  Sizes come from the environment:
    EAT_SYNTH_CLIP_SECONDS (10)   EAT_SYNTH_TRAIN_CLIPS (4096)   EAT_SYNTH_TEST_CLIPS (1054)
The teacher files `ex_audioset.py` loads (`resources/passt_enemble_logits_mAP_495.npy`, `fname_to_index.pkl`,
ex_audioset.py:104-118) are written by `write_teacher_files`; a fraction of the clips is deliberately left out
of `fname_to_index` to exercise the unknown-teacher mask (ex_audioset.py:166-177).
"""
import os
import pickle
import zlib

import numpy as np
import torch
from torch.utils.data import Dataset as TorchDataset, WeightedRandomSampler

NUM_CLASSES = 527
dataset_config = {"num_of_classes": NUM_CLASSES}


def _env_int(name, default):
    return int(os.environ.get(name, default))


def _seed(tag, i):
    return (zlib.crc32(f"{tag}{i}".encode()) ^ 0x5EA7) & 0x7FFFFFFF


def clip_name(split, i):
    return f"synth_{split}_{i:07d}"


def synth_target(split, i, n_clips):
    """multi-hot labels: class (i mod 527) and its successor chain so that every class is hit, plus ~2 random ones"""
    y = np.zeros(NUM_CLASSES, dtype=np.float32)
    y[i % NUM_CLASSES] = 1.0
    rng = np.random.RandomState(_seed(split + "y", i))
    y[rng.randint(0, NUM_CLASSES, size=2)] = 1.0
    return y


def synth_clip(split, i, n_samples, resample_rate=32000):
    g = torch.Generator()
    g.manual_seed(_seed(split + "x", i))
    x = torch.empty(n_samples, dtype=torch.float32).normal_(0.0, 0.1, generator=g)
    t = torch.arange(n_samples, dtype=torch.float32) / float(resample_rate)
    f0 = 100.0 + 25.0 * (i % NUM_CLASSES)                       # the first label is audible: a tone per class
    x += 0.05 * torch.sin(2 * np.pi * f0 * t) + 0.02 * torch.sin(2 * np.pi * (2.5 * f0 + 31.0) * t)
    return x.numpy()


class SyntheticAudioSet(TorchDataset):
    def __init__(self, split, n_clips, resample_rate=32000, clip_seconds=None, gain_augment=0):
        self.split, self.n, self.resample_rate = split, n_clips, resample_rate
        secs = float(os.environ.get("EAT_SYNTH_CLIP_SECONDS", 10)) if clip_seconds is None else clip_seconds
        self.n_samples = int(round(secs * resample_rate))
        self.gain_augment = gain_augment

    def __len__(self):
        return self.n

    def __getitem__(self, index):
        x = synth_clip(self.split, index, self.n_samples, self.resample_rate)
        if self.gain_augment:                                   # datasets/audioset.py:56-61
            gain = torch.randint(self.gain_augment * 2, (1,)).item() - self.gain_augment
            x = x * (10 ** (gain / 20))
        return x.reshape(1, -1), clip_name(self.split, index), synth_target(self.split, index, self.n)


class AddIndexDataset(TorchDataset):
    def __init__(self, ds):
        self.ds = ds

    def __getitem__(self, index):
        x, f, y = self.ds[index]
        return x, f, y, index

    def __len__(self):
        return len(self.ds)


def get_full_training_set(add_index=True, roll=False, wavmix=False, gain_augment=0, resample_rate=32000):
    if roll or wavmix:
        raise NotImplementedError("the synthetic AudioSet stand-in has no roll / waveform-mixup augmentation")
    ds = SyntheticAudioSet("train", _env_int("EAT_SYNTH_TRAIN_CLIPS", 4096), resample_rate, gain_augment=gain_augment)
    return AddIndexDataset(ds) if add_index else ds


def get_test_set(resample_rate=32000):
    return SyntheticAudioSet("eval", _env_int("EAT_SYNTH_TEST_CLIPS", 2 * NUM_CLASSES), resample_rate)


def get_ft_weighted_sampler(epoch_len=100000, sampler_replace=False):
    n = _env_int("EAT_SYNTH_TRAIN_CLIPS", 4096)
    all_y = torch.from_numpy(np.stack([synth_target("train", i, n) for i in range(n)]))
    per_class = 100.0 + all_y.sum(0).reshape(1, -1)             # datasets/audioset.py:199-210
    weights = (all_y * (1000.0 / per_class)).sum(dim=1)
    return WeightedRandomSampler(weights, num_samples=min(epoch_len, n) if not sampler_replace else epoch_len,
                                 replacement=sampler_replace)


def write_teacher_files(resources_dir, unknown_every=5, seed=0):
    """resources/passt_enemble_logits_mAP_495.npy ([n_known, 527] teacher LOGITS) and resources/fname_to_index.pkl
    ({audio_name: row}); every `unknown_every`-th training clip has no entry."""
    os.makedirs(resources_dir, exist_ok=True)
    n = _env_int("EAT_SYNTH_TRAIN_CLIPS", 4096)
    known = [i for i in range(n) if unknown_every <= 0 or i % unknown_every != unknown_every - 1]
    rng = np.random.RandomState(seed)
    logits = rng.normal(-3.0, 1.5, size=(len(known), NUM_CLASSES)).astype(np.float32)
    for row, i in enumerate(known):
        logits[row] += 5.0 * synth_target("train", i, n)        # a teacher that mostly agrees with the labels
    np.save(os.path.join(resources_dir, "passt_enemble_logits_mAP_495.npy"), logits)
    with open(os.path.join(resources_dir, "fname_to_index.pkl"), "wb") as f:
        pickle.dump({clip_name("train", i): row for row, i in enumerate(known)}, f)
    return len(known), n
