"""Test / validation runs of a trained checkpoint over a dataset split -- the B200 counterpart of the reference's
run_pipeline.py (:13-60 CLI, :103-150 config -> dataset / model / pipeline, :152-157 split dispatch):

  python run_pipeline.py --cfg_file configs/Liquid3d.yml --dataset_path <dir> --ckpt_path checkpoints/Liquid3d/ckpt \\
      --split test --output_dir ./output

``--split test`` rolls out every sequence of <dataset_path>/test and writes <output_dir>/visual/<seq>/<epoch>.npz (pred / gt /
bnd, the reference's dataset names); ``--split valid`` prints the metric means of run_valid.  ``--split train`` is out of
scope: the training STEP exists (dmcf_b200/training.py), the curriculum around it does not."""
import argparse
import json
import os
import logging
import sys


def main(argv=None):
    from dmcf_b200 import config, pipeline
    from dmcf_b200.simulator import Simulator
    ap = argparse.ArgumentParser(description="Run a trained DMCF network over a dataset split on a B200")
    ap.add_argument("-c", "--cfg_file", required=True, help="path to the config file")
    ap.add_argument("--dataset_path", help="path to the dataset (overrides dataset.dataset_path)")
    ap.add_argument("--ckpt_path", help="path to the checkpoint (overrides model.ckpt_path)")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--split", default="test", choices=["train", "valid", "test"])
    ap.add_argument("--output_dir", help="the dir to save outputs (overrides pipeline.output_dir)")
    args, extra = ap.parse_known_args(argv)
    if args.split == "train":
        raise NotImplementedError("run_train is out of scope (DESIGN.md section 8); see dmcf_b200/training.py for the step")
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    cfg = config.load_config(args.cfg_file, config.parse_cli_overrides(extra))
    if args.dataset_path:
        cfg["dataset"]["dataset_path"] = args.dataset_path
    model_cfg = dict(cfg["model"])
    ckpt = args.ckpt_path or model_cfg.pop("ckpt_path", None)
    model_cfg.pop("ckpt_path", None)
    model = config.build_model(model_cfg)
    sim = Simulator(model, device=args.device)
    epoch = sim.load_ckpt(ckpt)
    dataset = pipeline.open_split(cfg["dataset"], args.split)
    out_dir = args.output_dir or cfg["pipeline"].get("output_dir", "./output")
    if args.split == "test":
        valid_ds = None
        if cfg["pipeline"].get("test_compute_metric", False):  # the reference's run_valid rolls out dataset.valid
            vdir = os.path.join(cfg["dataset"].get("dataset_path") or "", "valid")
            valid_ds = pipeline.open_split(cfg["dataset"], "valid") if os.path.isdir(vdir) else None
        written, valid = pipeline.run_test(sim, dataset, cfg["pipeline"], out_dir, epoch, valid_dataset=valid_ds)
        print("\n".join(written))
        if valid is not None:
            print(json.dumps(valid))
    else:
        print(json.dumps(pipeline.run_valid(sim, dataset, cfg["pipeline"], epoch)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
