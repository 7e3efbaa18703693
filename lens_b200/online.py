"""Online / streaming matcher of the event-driven deployment (lens/run_speck.py:155-226) on the GPU.

The reference collects the output-layer spikes of every readout interval (`custom_readout`,
run_speck.py:155-174), and its `seq_match` thread (run_speck.py:177-226) turns every 4 readouts into
one sequence row (`vector // 4`, where `vector` holds the spike counts accumulated SINCE THE LAST
MATCH, not since the last row), and every 4 rows into one match:
`convolve2d(sequence.T, eye(L), mode='same') / L` followed by a per-column argmax; results are
concatenated into `matrix` whose transpose is saved as similarity_matrix.npy.

`OnlineMatcher.push(counts)` is one readout; the arithmetic happens in lens_online_accumulate /
lens_online_match (include/lens_b200.h).  The spike counts come from the GPU network
(`B200Network.run_streams`) instead of the Speck chip.
"""
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr, require_cuda


class OnlineMatcher:
    READOUTS_PER_ROW = 4      # run_speck.py:180  `if self.qry == 4`
    ROWS_PER_MATCH = 4        # run_speck.py:200  `if self.sequence.shape[0] == 4`

    def __init__(self, reference_places, sequence_length, device="cuda",
                 readouts_per_row=READOUTS_PER_ROW, rows_per_match=ROWS_PER_MATCH):
        if sequence_length < 1:
            raise ValueError("sequence_length must be >= 1")
        self.reference_places = int(reference_places)
        self.sequence_length = int(sequence_length)
        self.readouts_per_row = int(readouts_per_row)
        self.rows_per_match = int(rows_per_match)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.LensError("OnlineMatcher needs a CUDA device (there is no CPU path)")
        P, R = self.reference_places, self.rows_per_match
        self.sum = torch.zeros(P, dtype=torch.int32, device=self.device)            # run_speck.py `self.sum`
        self.sequence = torch.zeros((R, P), dtype=torch.int32, device=self.device)  # `self.sequence`
        self.qry = 0                                                                 # `self.qry`
        self.n_rows = 0
        self.matrix = None          # [P, 4 * matches] float64, like the reference's `self.matrix`
        self.matches = []           # argmax per column of every match

    def reset(self):
        self.sum.zero_()
        self.qry = 0
        self.n_rows = 0
        self.matrix = None
        self.matches = []

    def push(self, counts):
        """One readout: counts f32 [P] (spikes per place in this interval).

        Returns None, or (argmax i32 [rows_per_match], result f64 [P, rows_per_match]) when this readout
        completes a sequence (run_speck.py:200-222)."""
        require_cuda(counts)
        if counts.dtype != torch.float32 or counts.numel() != self.reference_places:
            raise ValueError("counts must be float32 [reference_places]")
        counts = counts.contiguous()
        self.qry += 1
        row_done = self.qry == self.readouts_per_row
        row = self.sequence[self.n_rows] if row_done else None
        check(_lib.lib().lens_online_accumulate(ptr(self.sum), ptr(counts), self.reference_places,
                                                self.readouts_per_row, ptr(row), stream_ptr()),
              "lens_online_accumulate")
        if not row_done:
            return None
        self.qry = 0
        self.n_rows += 1
        if self.n_rows < self.rows_per_match:
            return None
        P, R = self.reference_places, self.rows_per_match
        result = torch.empty((P, R), dtype=torch.float64, device=self.device)
        argmax = torch.empty(R, dtype=torch.int32, device=self.device)
        check(_lib.lib().lens_online_match(ptr(self.sequence), R, P, self.sequence_length, ptr(result), ptr(argmax),
                                           stream_ptr()), "lens_online_match")
        self.matrix = result if self.matrix is None else torch.cat((self.matrix, result), dim=1)
        self.matches.append(argmax)
        self.sum.zero_()            # run_speck.py:221-222
        self.n_rows = 0
        return argmax, result

    def similarity_matrix(self):
        """What the reference saves as similarity_matrix.npy (run_speck.py:219): matrix.T."""
        return None if self.matrix is None else self.matrix.T
