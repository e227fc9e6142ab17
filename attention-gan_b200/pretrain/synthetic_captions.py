"""Seeded synthetic stand-in for the reference's caption pipeline (data/bedrooms.py:241-304).

The reference has no text: ``HierarchicalClusterer.cluster`` clusters image embeddings at k = 7, 15, 31, 62, 125,
250, 500 (``_determine_k_values``: max_vocab_size // 2^j while > min_clusters, ascending; bedrooms.py:292-304 with the
arguments of pretrain_damsm.py:57) and the "caption" of an image is the list of its cluster labels from coarse to
fine, ``['k7c3', 'k15c9', ..., 'k500c231']``: FIXED length 7, vocabulary <= 990 words, class id = the finest
cluster (bedrooms.py:262-268).  Complete-linkage cuts of one dendrogram are nested, so a finer cluster determines
all coarser ones.  This module generates captions with exactly that structure from a seed, without images, UMAP or
sklearn: a random nested hierarchy over the k values, one leaf per sample.

Word indices follow ``Vocab._addWord`` (bedrooms.py:94-99): words are numbered in order of first occurrence over the
dataset, there are no special tokens.
"""
from __future__ import annotations

from typing import List

import numpy as np
import torch


def hierarchical_k_values(max_vocab_size: int = 1000, min_k: int = 5) -> List[int]:
    """bedrooms.py:292-304 (`_determine_k_values`): ascending cluster counts"""
    factor, out = 2, []
    k = max_vocab_size // factor
    while k > min_k:
        out.append(k)
        factor *= 2
        k = max_vocab_size // factor
    return list(reversed(out))


class SyntheticBedroomCaptions:
    """captions [N, 7] int64 (word indices), lengths [N] (all 7), class_ids [N] (finest cluster), vocab_size"""

    def __init__(self, n_images: int, max_vocab_size: int = 1000, min_clusters: int = 5, seed: int = 0):
        rng = np.random.default_rng(seed)
        self.k_values = hierarchical_k_values(max_vocab_size, min_clusters)
        ks = self.k_values
        # nested hierarchy: parent[l][c] = cluster at level l-1 that contains cluster c of level l; every coarse
        # cluster gets at least one child (a cut of a dendrogram never leaves a cluster empty)
        self.parent = [None]
        for l in range(1, len(ks)):
            par = np.concatenate([np.arange(ks[l - 1]), rng.integers(0, ks[l - 1], ks[l] - ks[l - 1])])
            rng.shuffle(par)
            self.parent.append(par)
        leaf = rng.integers(0, ks[-1], n_images)                 # the finest cluster of every image
        labels = np.zeros((n_images, len(ks)), np.int64)
        labels[:, -1] = leaf
        for l in range(len(ks) - 1, 0, -1):
            labels[:, l - 1] = self.parent[l][labels[:, l]]
        self.cluster_labels = labels                              # [N, levels] cluster index per level, coarse -> fine
        # vocabulary in order of first occurrence of 'k{k}c{c}' scanning captions left to right (Vocab._addWord)
        word2index = {}
        caps = np.zeros_like(labels)
        for i in range(n_images):
            for l, k in enumerate(ks):
                w = (k, int(labels[i, l]))
                if w not in word2index:
                    word2index[w] = len(word2index)
                caps[i, l] = word2index[w]
        self.word2index = {f"k{k}c{c}": v for (k, c), v in word2index.items()}
        self.vocab_size = len(word2index)
        # class ids: enumerate the distinct finest clusters (bedrooms.py:262-268; the reference enumerates a set)
        uniq = {c: j for j, c in enumerate(sorted(set(leaf.tolist())))}
        self.captions = torch.from_numpy(caps)
        self.lengths = torch.full((n_images,), len(ks), dtype=torch.int64)
        self.class_ids = torch.tensor([uniq[c] for c in leaf.tolist()], dtype=torch.int64)

    def __len__(self) -> int:
        return self.captions.shape[0]

    def batch(self, start: int, size: int):
        """(captions [size,7], lengths [size], class_ids [size]) like one DataLoader batch (bedrooms.py:229-238)"""
        sl = slice(start, start + size)
        return self.captions[sl], self.lengths[sl], self.class_ids[sl]
