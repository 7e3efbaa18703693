"""Inference driver with the reference's Python API (lens/run_model.py).

    model = LENS(args)                    # args: the reference's argparse Namespace (main.py:82-189)
    R = run_inference(model, model_name)  # -> [Recall@1, @5, @10, @15, @20, @25] rounded to 2 dp

What changes underneath: the converted sinabs network is a `B200Network` (hand-written sm_100a
kernels behind the C ABI), the per-query python loop of run_model.py:229-246 is one batched
`run_streams` call, sequence matching + top-N + the Recall@N counters run on the GPU
(lens_seqmatch_topk / lens_recall), and only ground-truth preparation (run_model.py:266-294,
a one-off scipy dilation) and plotting stay on the host.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import LensError
from .network import B200Network
from .src import blitnet as bn
from .src.dataset import CustomImageDataset, ProcessImage
from .src.loggers import model_logger
from .src.metrics import recallAtK, createPR, recall_detail  # noqa: F401  (re-exported like the reference module)

RECALL_NS = [1, 5, 10, 15, 20, 25]   # run_model.py:266


def create_GTtol(GT, distance=2):
    """Ground truth with tolerance: binary dilation by a (2*distance+1)^2 square (run_model.py:272-289)."""
    from scipy.ndimage import binary_dilation
    se = np.ones((2 * distance + 1, 2 * distance + 1), dtype=int)
    return binary_dilation(GT, structure=se).astype(int)


class LENS(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.args = args
        for arg in vars(args):
            setattr(self, arg, getattr(args, arg))
        self.dataset_file = os.path.join(self.data_dir, self.query + ".csv")
        self.query_dir = os.path.join(self.data_dir, self.dataset, self.camera, self.query)
        self.reference_dir = os.path.join(self.data_dir, self.dataset, self.camera, self.reference)

        self.device = model_logger(self)
        if getattr(self, "nocuda", False):
            self.logger.warning("--nocuda ignored: lens_b200 has no CPU path, running on the B200")
        if getattr(self, "simulated_speck", False):
            raise LensError("--simulated_speck needs Speck2fDevKit hardware (out of scope for lens_b200)")

        self.layer_dict = {}
        self.layer_counter = 0
        self.input = int(args.dims * args.dims)
        self.feature = int(self.input * self.feature_multiplier)
        self.output = int(args.reference_places)
        self.add_layer("feature_layer", dims=[self.input, self.feature], device=self.device, inference=True)
        self.add_layer("output_layer", dims=[self.feature, self.output], device=self.device, inference=True)
        if not hasattr(self, "matrix"):
            self.matrix = None
        self.kernel_size = self.roi_dim // self.dims
        self.sinabs_model = None
        self.similarity = None        # [Q, P] spike counts of the last evaluate()
        self.dist_matrix_seq = None   # sequence-matched matrix of the last evaluate()
        self.recall_gpu = None        # Recall@N under the CUDA kernel's deterministic tie rule
        self.recall_bounds = None     # [(lo, hi)] over every order of equal similarities

    def add_layer(self, name, **kwargs):
        if name in self.layer_dict:
            raise ValueError(f"Layer with name {name} already exists.")
        setattr(self, name, bn.SNNLayer(**kwargs))
        self.layer_dict[name] = self.layer_counter
        self.layer_counter += 1

    # ------------------------------------------------------------------------------
    def build_network(self, max_streams=1):
        """What run_model.py:130-156 assembles: pooling conv + IAF + Linear + IAF + Linear + IAF."""
        self.sinabs_model = B200Network(self.feature_layer.w.weight, self.output_layer.w.weight,
                                        roi=self.roi_dim, k=self.kernel_size,
                                        num_timesteps=self.timebin, max_streams=max_streams,
                                        device=self.device)
        if self.sinabs_model.n_inexact:
            self.logger.warning(f"{self.sinabs_model.n_inexact} weights lie more than 22 bits below "
                                "their row maximum and were rounded to the fixed-point grid")
        return self.sinabs_model

    def similarity_matrix(self, test_loader):
        """Rows of spike counts, one per query (run_model.py:229-246) -> f32 [Q, P] on the GPU."""
        ds = getattr(test_loader, "dataset", test_loader)
        if hasattr(ds, "load_frames"):
            frames = ds.load_frames().to(self.device)               # u8 [Q, roi, roi]
            return self.sinabs_model.run_streams(frames=frames[None])[0]
        # generic loader yielding the reference's (spikes [1, T, 1, roi, roi], label, _, _) tuples
        rows = []
        for spikes, _, _, _ in test_loader:
            spikes = spikes.to(self.device)
            spikes = spikes.reshape(-1, *spikes.shape[2:])          # sl.FlattenTime()
            rows.append(self.sinabs_model(spikes).sum(dim=0).reshape(-1))
        return torch.stack(rows)

    def evaluate(self, test_loader, model):
        self.build_network(max_streams=1)
        with torch.no_grad():
            S = self.similarity_matrix(test_loader)                 # [Q, P] f32, device
        self.sinabs_model.raise_on_overflow()     # > 127 spikes of one neuron in one step: clipped, not the reference's
        Q, P = S.shape
        if (Q, P) != (model.query_places, model.reference_places):
            raise LensError(f"similarity matrix is {Q}x{P}, expected "
                            f"{model.query_places}x{model.reference_places}")
        self.similarity = S
        L = int(self.sequence_length)
        n_top = max(RECALL_NS)
        if L != 0:
            _, top_idx, D = ops.seqmatch_topk(S[None].contiguous(), L, n_top, want_D=True)
            D_dev = D[0]
            dist_matrix_seq = D_dev.cpu().numpy()                   # [P-L+1, Q-L+1], rows = database
        else:
            # run_model.py:254 keeps `out` itself (rows = query) and ranks along axis 0
            _, top_idx, _ = ops.seqmatch_topk(S.t().contiguous()[None], 1, n_top)
            D_dev = S
            dist_matrix_seq = S.cpu().numpy().astype(np.float64)
        self.dist_matrix_seq = dist_matrix_seq
        self._save_matrix_pdf(dist_matrix_seq, "distance_matrix_lens.pdf")

        R = []
        GTtol = None
        if self.matching:
            GT = np.load(os.path.join(self.data_dir, self.dataset, self.camera,
                                      self.reference + "_" + self.query + "_GT.npy"))
            if L != 0:
                GT = GT[L - 2:-1, L - 2:-1]
            GTtol = create_GTtol(GT, distance=self.GT_tolerance).T
            self._save_matrix_pdf(GTtol, "GTtol.pdf")
            # the reference fails inside recallAtK with this very assertion (metrics.py:196), e.g. for
            # --sequence_length 1, whose GT slice GT[-1:-1] is empty
            assert GTtol.shape == dist_matrix_seq.shape, "S_in and GThard must have the same shape"
            # Recall@N on the GPU: deterministic tie rule + bounds over every tie order.  Where the bounds
            # differ, the reference's number depends on how numpy's unstable argsort orders equal
            # similarities (metrics.py:218); that order is then taken from the same numpy call.
            det = recall_detail(dist_matrix_seq, GTtol.astype(bool), RECALL_NS, top_idx=top_idx, S_dev=D_dev)
            self.recall_gpu = [round(r, 2) for r in det["gpu"]]
            self.recall_bounds = [(round(a, 2), round(b, 2)) for a, b in zip(det["lo"], det["hi"])]
            R = [round(r, 2) for r in det["reference"]]
            model.logger.info("N      " + "  ".join(f"{n:>5d}" for n in RECALL_NS))
            model.logger.info("Recall " + "  ".join(f"{r:>5.2f}" for r in R))
            if det["ties_decided"]:
                model.logger.info("ties decide Recall@%s: GPU tie rule %s, bounds %s" %
                                  (det["ties_decided"], self.recall_gpu, self.recall_bounds))
        self.GTtol = GTtol

        if getattr(self, "PR_curve", False):
            if GTtol is None:
                raise LensError("--PR_curve needs --matching (the reference fails the same way: GTtol undefined)")
            LENS_P, LENS_R = createPR(dist_matrix_seq.T, GTtol.T, self.output_folder, matching="single", n_thresh=100)
            self.lens_PR = {"Precision": LENS_P, "Recall": LENS_R}
        if getattr(self, "sad", False):
            if GTtol is None:
                raise LensError("--sad needs --matching (the reference fails the same way: GTtol undefined)")
            from .src.sad import run_sad
            self.sad_PR, self.sad_Recall = run_sad(self.reference_dir, self.query_dir, GTtol, self.output_folder,
                                                   self.sequence_length)
            model.logger.info("SAD    " + "  ".join(f"{r:>5.2f}" for r in self.sad_Recall))

        model.logger.info("")
        model.logger.info("Succesfully completed inferencing using LENS")
        return R

    def _save_matrix_pdf(self, mat, name):
        """Optional plot (run_model.py:256-260,296-299); skipped when matplotlib is not installed."""
        try:
            import matplotlib
            matplotlib.use("Agg")
            import matplotlib.pyplot as plt
        except ImportError:
            return
        plt.imshow(mat)
        plt.colorbar()
        plt.savefig(os.path.join(self.output_folder, name))
        plt.close()

    def forward(self, spikes):
        return self.sinabs_model(spikes)

    def load_model(self, model_path):
        self.load_state_dict(torch.load(model_path, map_location=self.device, weights_only=True),
                             strict=False)


def run_inference(model, model_name, models_dir="./lens/models"):
    """lens/run_model.py:360-397: build the query dataset, load `models_dir/model_name`, evaluate."""
    test_dataset = CustomImageDataset(annotations_file=model.dataset_file, img_dir=model.query_dir,
                                      transform=ProcessImage(), kernel_size=model.kernel_size,
                                      skip=model.filter, max_samples=model.query_places,
                                      is_spiking=True, time_window=model.timebin)
    model.eval()
    model.load_model(os.path.join(models_dir, model_name))
    return model.evaluate(test_dataset, model)
