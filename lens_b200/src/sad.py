"""Sum-of-absolute-differences baseline with the reference's interface (lens/src/sad.py).

`run_sad(reference, query, GT, outputdir, sequence_length) -> (PR_data, recallatn)`: every stage is a
CUDA kernel of liblens_b200.so -- the L1 distance matrix (lens_sad_matrix), the diagonal sequence
matching (lens_seqmatch_topk), 1 / distance (lens_reciprocal), createPR and Recall@N.  Only the PNG
decoding and the natural sort of file names stay on the host; the plots of the reference are omitted.
"""
import os
import re

import torch

from .. import ops
from .dataset import read_png_u8
from .metrics import createPR, recallAtK

RECALL_NS = [1, 5, 10, 15, 20, 25]


def natural_sort_key(s):
    return [int(t) if t.isdigit() else t.lower() for t in re.split("([0-9]+)", s)]


def load_and_preprocess_images(folder_path):
    """All .png files of a folder in natural order, flattened: u8 [n, H*W] (lens/src/sad.py:15-23)."""
    files = sorted(os.listdir(folder_path), key=natural_sort_key)
    imgs = [read_png_u8(os.path.join(folder_path, f)).reshape(-1) for f in files if f.endswith(".png")]
    return torch.stack(imgs)


def sad_distance_matrix(query_frames, reference_frames, sequence_length):
    """u8 frames -> sequence-matched distance matrix f32 [R-L+1, Q-L+1] on the GPU (sad.py:38-42)."""
    dist = ops.sad_matrix(query_frames.cuda().contiguous(), reference_frames.cuda().contiguous())   # [Q, R]
    _, _, D = ops.seqmatch_topk(dist[None].contiguous(), sequence_length, 1, want_D=True)
    return D[0]


def run_sad(reference, query, GT, outputdir, sequence_length):
    images1 = load_and_preprocess_images(query)
    images2 = load_and_preprocess_images(reference)
    D = sad_distance_matrix(images1, images2, sequence_length)
    sim = ops.reciprocal(D.contiguous()).cpu().numpy()            # 1 / dist_matrix_seq
    P, R = createPR(sim, GT, outputdir, datatype="SAD", matching="single", n_thresh=100)
    PR_data = {"Precision": P, "Recall": R}
    recallatn = [round(recallAtK(sim, GT, K=n), 2) for n in RECALL_NS]
    return PR_data, recallatn
