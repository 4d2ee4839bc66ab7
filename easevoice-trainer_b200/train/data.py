"""Stage-2 input side: dataset over the files Normalize writes, pad-collate and the length-bucketed distributed
sampler (reference: src/easevoice/module/data_utils.py:14-324).  Differences: the wav files under 5-wav32k are read
directly (they are already 32 kHz mono int16; the reference pipes each one through an ffmpeg subprocess), and |X| is
NOT computed per item on CPU workers -- the trainer computes it for the whole batch on the GPU with the fused mel
kernel (s2_step.to_device_batch), which takes the per-item sample counts (`wav_lengths`) so that every row is
reflect-padded at its own end and zero beyond its own last frame, exactly like the reference's per-item features.
"""
import math
import os
import random
import wave

import numpy as np
import torch


def default_phoneme_table():
    """phoneme -> id map.  The 732-entry SYMBOLS vocabulary belongs to the host application (reference:
    src/easevoice/text/symbols.py:410-412); as a drop-in this package imports it from there."""
    from src.easevoice.text.symbols import SYMBOLS          # noqa: provided by the reference application
    return {s: i for i, s in enumerate(SYMBOLS)}


class TextAudioSpeakerLoader(torch.utils.data.Dataset):
    """data_utils.py:14-117 (selection / filtering logic identical; items are (ssl, wav, text))."""

    def __init__(self, exp_dir, sampling_rate=32000, hop_length=640, phoneme_table=None, val=False):
        self.path2 = f"{exp_dir}/2-name2text.txt"
        self.path4 = f"{exp_dir}/4-cnhubert"
        self.path5 = f"{exp_dir}/5-wav32k"
        for p in (self.path2, self.path4, self.path5):
            assert os.path.exists(p), p
        table = phoneme_table or default_phoneme_table()
        names4 = {n[:-3] for n in os.listdir(self.path4)}
        names5 = set(os.listdir(self.path5))
        phon = {}
        with open(self.path2, encoding="utf8") as f:
            for line in f.read().strip("\n").split("\n"):
                t = line.split("\t")
                if len(t) == 4:
                    phon[t[0]] = t[1]
        names = list(set(phon) & names4 & names5)
        if len(names) < 100:
            names = names * max(2, int(100 / max(len(names), 1)))
        random.seed(1234)
        random.shuffle(names)
        self.items, self.lengths = [], []
        self.sampling_rate, self.hop_length = sampling_rate, hop_length
        for n in names:
            try:
                ids = [table[p] for p in phon[n].split(" ")]
            except KeyError:
                continue
            size = os.path.getsize(f"{self.path5}/{n}")
            dur = size / sampling_rate / 2
            if dur == 0 or not (54 > dur > 0.6 or val):
                continue
            self.items.append((n, ids))
            self.lengths.append(size // (2 * hop_length))
        if len(self.items) <= 1:
            raise ValueError(f"data in {exp_dir} is all skipped, please check the data")

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        name, ids = self.items[i]
        with wave.open(f"{self.path5}/{name}", "rb") as w:
            pcm = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
        wav = torch.from_numpy(pcm.astype(np.float32) / 32768.0).unsqueeze(0)
        ssl = torch.load(f"{self.path4}/{name}.pt", map_location="cpu").float()
        T = wav.shape[1] // self.hop_length
        if ssl.shape[-1] != T:                                   # data_utils.py:101-104
            ssl = torch.nn.functional.pad(ssl, (0, 1), mode="replicate")
        return ssl, wav, torch.tensor(ids, dtype=torch.long)


class TextAudioSpeakerCollate:
    """data_utils.py:167-226: zero-pad to the batch maximum, sorted by length (descending); frame axis padded to
    2*(Tmax//2+1).  Returns the host-side dict consumed by s2_step.to_device_batch."""

    def __init__(self, hop_length=640):
        self.hop = hop_length

    def __call__(self, batch):
        order = sorted(range(len(batch)), key=lambda i: batch[i][1].shape[1] // self.hop, reverse=True)
        T = max(b[1].shape[1] // self.hop for b in batch)
        Tp = 2 * (T // 2 + 1)
        L = Tp * self.hop
        X = max(b[2].shape[0] for b in batch)
        B = len(batch)
        ssl = torch.zeros(B, 768, Tp)
        wav = torch.zeros(B, 1, L)
        text = torch.zeros(B, X, dtype=torch.long)
        lengths = torch.zeros(B, dtype=torch.long)
        wav_lengths = torch.zeros(B, dtype=torch.long)
        text_lengths = torch.zeros(B, dtype=torch.long)
        for k, i in enumerate(order):
            s, w, t = batch[i]
            n = w.shape[1] // self.hop
            ssl[k, :, :min(s.shape[-1], Tp)] = s[0, :, :Tp]
            wav[k, :, :w.shape[1]] = w
            text[k, :t.shape[0]] = t
            lengths[k], text_lengths[k], wav_lengths[k] = n, t.shape[0], w.shape[1]
        return dict(ssl=ssl, wav=wav, text=text, lengths=lengths, text_lengths=text_lengths, wav_lengths=wav_lengths)


class DistributedBucketSampler(torch.utils.data.Sampler):
    """data_utils.py:229-324: batches of similar length; rank r takes ids_bucket[r::num_replicas] of every bucket."""

    def __init__(self, lengths, batch_size, boundaries, num_replicas=1, rank=0, shuffle=True):
        self.lengths, self.batch_size, self.boundaries = list(lengths), batch_size, list(boundaries)
        self.num_replicas, self.rank, self.shuffle, self.epoch = num_replicas, rank, shuffle, 0
        buckets = [[] for _ in range(len(self.boundaries) - 1)]
        for i, ln in enumerate(self.lengths):
            k = self._bisect(ln)
            if k != -1:
                buckets[k].append(i)
        for i in range(len(buckets) - 1, -1, -1):
            if not buckets[i]:
                buckets.pop(i)
                self.boundaries.pop(i + 1)
        self.buckets = buckets
        total = num_replicas * batch_size
        self.num_samples_per_bucket = [len(b) + (total - len(b) % total) % total for b in buckets]
        self.total_size = sum(self.num_samples_per_bucket)
        self.num_samples = self.total_size // num_replicas

    def set_epoch(self, epoch):
        self.epoch = epoch

    def _bisect(self, x):
        for i in range(len(self.boundaries) - 1):
            if self.boundaries[i] < x <= self.boundaries[i + 1]:
                return i
        return -1

    def __iter__(self):
        g = torch.Generator()
        g.manual_seed(self.epoch)
        batches = []
        for bucket, nsb in zip(self.buckets, self.num_samples_per_bucket):
            ids = torch.randperm(len(bucket), generator=g).tolist() if self.shuffle else list(range(len(bucket)))
            rem = nsb - len(bucket)
            ids = ids + ids * (rem // len(bucket)) + ids[:rem % len(bucket)]
            ids = ids[self.rank::self.num_replicas]
            for j in range(len(ids) // self.batch_size):
                batches.append([bucket[k] for k in ids[j * self.batch_size:(j + 1) * self.batch_size]])
        if self.shuffle:
            batches = [batches[i] for i in torch.randperm(len(batches), generator=g).tolist()]
        assert len(batches) * self.batch_size == self.num_samples
        return iter(batches)

    def __len__(self):
        return self.num_samples // self.batch_size
