"""The reference's configuration surface (main.py:82-189) for the inference path."""
import argparse


def build_parser():
    p = argparse.ArgumentParser(description="LENS inference on B200 (flags of the reference's main.py)")
    p.add_argument("--dataset", type=str, default="example")
    p.add_argument("--camera", type=str, default="davis128")
    p.add_argument("--data_name", type=str, default="experiment001")
    p.add_argument("--reference", type=str, default="example-reference")
    p.add_argument("--query", type=str, default="example-query")
    p.add_argument("--data_dir", type=str, default="./lens/dataset/")
    p.add_argument("--reference_places", type=int, default=100)
    p.add_argument("--query_places", type=int, default=100)
    p.add_argument("--sequence_length", type=int, default=2)
    p.add_argument("--feature_multiplier", type=float, default=2.0)
    p.add_argument("--filter", type=int, default=1)
    p.add_argument("--dims", type=int, default=10)
    p.add_argument("--roi_dim", type=int, default=80)
    p.add_argument("--GT_tolerance", type=int, default=3)
    p.add_argument("--timebin", type=int, default=250)
    # training hyper-parameters (main.py:113-153): accepted so that a reference command line parses; the
    # inference path copies them onto the model like the reference (run_model.py:59-60) and never reads them
    p.add_argument("--epoch_feat", type=int, default=128)
    p.add_argument("--epoch_out", type=int, default=128)
    for flag, default in (("thr_l_feat", 0.0), ("thr_h_feat", 0.75), ("fire_l_feat", 0.4), ("fire_h_feat", 0.6),
                          ("ip_rate_feat", 0.02), ("stdp_rate_feat", 0.01), ("thr_l_out", 0.0),
                          ("thr_h_out", 0.5), ("fire_l_out", 0.5), ("fire_h_out", 0.5), ("ip_rate_out", 0.02),
                          ("stdp_rate_out", 0.01), ("f_exc", 0.35), ("f_inh", 0.75), ("o_exc", 1.0),
                          ("o_inh", 1.0)):
        p.add_argument("--" + flag, type=float, default=default)
    for flag in ("train_model", "sim_mat", "PR_curve", "matching", "sad", "nocuda", "event_driven",
                 "simulated_speck", "collect_data", "headless", "save_input"):
        p.add_argument("--" + flag, action="store_true")
    return p


def default_args(**overrides):
    args = build_parser().parse_args([])
    for k, v in overrides.items():
        setattr(args, k, v)
    return args


def generate_model_name(model):
    """main.py:27-38."""
    return ("".join(model.reference) + "_LENS_IN" + str(model.input) + "_FN" + str(model.feature) +
            "_DB" + str(model.reference_places) + ".pth")
