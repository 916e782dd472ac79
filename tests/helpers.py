"""Shared test helpers: rebuild the seeded synthetic sets that tests/golden/*.npz were made from."""
import importlib
import json
import os
from functools import lru_cache

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

synth = importlib.import_module("video-gen-evals_b200.synth")


def oracle():
    """The CPU oracle (test infrastructure; never imported by the product package)."""
    import sys
    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("oracle.tag_oracle")


def reference():
    """The unmodified reference modules from oracle/_ref (placed there by __graft_entry__.build() in the build container; the
    directory travels to the GPU box) or None."""
    import sys
    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    rr = importlib.import_module("oracle.ref_runner")
    return rr.load_ref()


def ref_runner():
    import sys
    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    return importlib.import_module("oracle.ref_runner")


class GoldenCase:
    def __init__(self, tag):
        self.tag = tag
        with open(os.path.join(GOLDEN, f"{tag}.json")) as f:
            self.meta = json.load(f)
        self.npz = np.load(os.path.join(GOLDEN, f"{tag}.npz"))
        m = self.meta
        self.appearance = m["appearance"]
        self.clip_len = m["clip_len"]
        self.stride = m["stride"]
        self.real = synth.make_videos(m["real_n"], m["real_len"], seed=m["real_seed"], appearance=self.appearance,
                                      name_prefix="real_")
        self.gen = synth.make_videos(len(m["gen_lens"]), m["gen_lens"], seed=m["gen_seed"],
                                     appearance=self.appearance, name_prefix="gen_")
        self.dims_raw, self.dims_diff = synth.dims_maps(self.appearance)
        self.mods = list(self.dims_raw.keys())
        self.sd = synth.make_state_dict(self.dims_raw, self.dims_diff, seed=m["seed_w"])
        self.label_dict = m["label_dict"]
        self.real_index = {n: i for i, n in enumerate(self.real.names)}
        self.gen_index = {n: i for i, n in enumerate(self.gen.names)}

    def stats(self):
        """golden ModalityStats as {'<prefix>_{raw,diff}_{mean,std}': tensor}"""
        out = {}
        for k in self.npz.files:
            if k.startswith("stats."):
                out[k[len("stats."):]] = torch.from_numpy(self.npz[k])
        return out

    def train_videos(self):
        return [self.real.video(self.real_index[n]) for n in self.meta["train_names"]]

    def gen_windows(self):
        return [(self.gen_index[n], s) for n, s in self.meta["gen_windows"]]

    def real_windows(self):
        return [(self.real_index[n], s) for n, s in self.meta["real_windows"]]


@lru_cache(maxsize=None)
def golden_case(tag):
    return GoldenCase(tag)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12)))


def max_abs(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))
