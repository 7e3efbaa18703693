"""`python -m lens_b200 [reference main.py flags]`: run the inference path on the GPU."""
from .config import build_parser, generate_model_name
from .run_model import LENS, run_inference


def main():
    args = build_parser().parse_args()
    if args.train_model or args.collect_data or args.event_driven:
        raise SystemExit("lens_b200 implements the inference path only (no training / Speck modes)")
    model = LENS(args)
    print(run_inference(model, generate_model_name(model)))


if __name__ == "__main__":
    main()
