"""B200-native TAG scoring hot path (drop-in for the scoring path of XThomasBU/video-gen-evals).

Public names mirror the reference modules they replace:
  model.py   -> HumanActionScorer
  utils.py   -> ModalityStats, WindowDataset, safe_collate, build_train_centroids_subset
  eval.py    -> infer_dims_from_stats, extract_window_features,
                compute_temporal_coherence_scores, compute_action_consistency_scores
  losses.py  -> TCL, SupConWithHardNegatives (forward); utils.py:65-95 hard-negative augmentations
plus the fused device-resident pipeline (`TagScorer`) used by bench.py.

Importing the package does not need a GPU; every compute entry point does, and raises `TagError`
when libtag_b200.so is missing (no CPU / PyTorch fallback exists).
"""
from ._lib import TagError, load as load_library, LIB_PATH
from .synth import (ACTION_CLASSES, VideoBatch, make_videos, make_state_dict, dims_maps, enumerate_windows,
                    sinusoidal_pe)
from .model import HumanActionScorer
from .features import (ModalityStats, DeviceVideos, FeatureFuser, WindowDataset, safe_collate, stats_vectors,
                       infer_dims_from_stats, compute_stats_from_videos)
from .scoring import (extract_window_features, compute_temporal_coherence_scores,
                      compute_action_consistency_scores, build_train_centroids_subset, centroid_accumulate,
                      centroid_finalize, allreduce_centroid_sums, write_video_scores)
from .losses import TCL, SupConWithHardNegatives, hard_negative_step
from .augment import partial_shuffle_within_window, reverse_sequence, get_static_window, gather_frames
from .pipeline import TagScorer, block_plan, shard_range, window_table
from .ingest import NpzIngest, FileItem, class_from_filename

__all__ = [n for n in dir() if not n.startswith("_")]
