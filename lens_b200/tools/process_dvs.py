"""Command line of the DVS tools (mirror of lens/tools/process_dvs.py:23-86): same flags and defaults.

    python -m lens_b200.tools.process_dvs --tool simple_rep --input_file sunset1 --reference ...
"""
import argparse
import sys

from .dvstools import FrameRep


def run_tool(args):
    if args.tool == "":
        print("No tool specified.")
        sys.exit()
    if args.tool in ("simple_rep", "decay_rep", "event_profile"):
        rep = FrameRep(args)
        for _ in rep.event_data():
            print("Processed")
        return rep
    raise SystemExit("tool '{}' is file-format conversion only (ROS bag / video) and is not part of lens_b200"
                     .format(args.tool))


def build_parser():
    p = argparse.ArgumentParser(description="Args for configuration DVS processing")
    p.add_argument("--tool", type=str, default="", help="Tool to implement")
    p.add_argument("--input_file", type=str, default="sunset1", help="Input file")
    p.add_argument("--hot_pixels", type=str, default="sunset1_hot_pixels", help="Hot pixels file")
    p.add_argument("--output_name", type=str, default="sunset1_profile", help="Output name")
    p.add_argument("--dataset_folder", type=str, default="./lens/dataset/brisbane_event/davis", help="Dataset folder")
    p.add_argument("--timebin", type=float, default=1, help="Timebin for frame representation (in fps)")
    p.add_argument("--decay_factor", type=float, default=5, help="Decay factor for frame representation")
    p.add_argument("--accum_factor", type=float, default=1, help="Accumulation value for frame representation")
    p.add_argument("--offset", type=float, default=1587452582.35, help="Offset for frame representation")
    p.add_argument("--frames_max", type=int, default=900, help="Number of frames to generate")
    p.add_argument("--frame_limit", action="store_true", help="Stop after frames_max frames")
    p.add_argument("--pixels", type=int, default=25, help="Number of patch centroids (output pixels)")
    p.add_argument("--reference", action="store_true", help="Draw and store a new patch layout")
    return p


def dvs_parser(argv=None):
    return run_tool(build_parser().parse_args(argv))


if __name__ == "__main__":
    dvs_parser()
