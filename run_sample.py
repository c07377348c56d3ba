"""Rollout of one frame file with a trained checkpoint -- the B200 counterpart of the reference's run_sample.py
(:13-60 CLI, :140-181 rollout with inflow, :200-235 data in / results out):

  python run_sample.py --cfg_file configs/Liquid3d.yml --ckpt_path checkpoints/Liquid3d/ckpt \\
      --data_path datasets/canyon_data/canyon.msgpack.zst --timesteps 800 --inflow 600 --output_dir ./output

Results go to <output_dir>/example/0000/<epoch>.npz with the reference's dataset names (<model>/pred, <model>/bnd)."""
import argparse
import os
import sys

import numpy as np
import torch


def run_rollout(sim, model, data, timesteps, inflow=0, inflow_velocity=(10.0, 0.0, -6.0)):
    """run_sample.py:140-181: the frame's velocities get the inflow offset, every second step (while t < inflow) the
    initial block of particles is appended again."""
    dev = sim.device
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    in_pos = t(data["pos"])
    in_vel = t(data["vel"]) + torch.tensor([inflow_velocity], dtype=torch.float32, device=dev)
    in_acc = torch.zeros_like(in_pos) + torch.tensor([[0.0, float(model.grav), 0.0]], dtype=torch.float32, device=dev)
    inputs = [in_pos, in_vel, in_acc, None, t(data["box"]), t(data["box_normals"])]
    results = [inputs[0]]
    for step in range(timesteps - 1):
        inputs = sim.run_inference([inputs])[0]
        results.append(inputs[0])
        if inflow > step and step % 2 == 1:
            inputs[0] = torch.cat([inputs[0], in_pos], dim=0)
            inputs[1] = torch.cat([inputs[1], in_vel], dim=0)
            inputs[2] = torch.cat([inputs[2], in_acc], dim=0)
    return results


def main(argv=None):
    from dmcf_b200 import config, datasets
    from dmcf_b200.simulator import Simulator
    ap = argparse.ArgumentParser(description="Rollout of a frame file with a trained DMCF checkpoint on a B200")
    ap.add_argument("-c", "--cfg_file", required=True, help="path to the config file")
    ap.add_argument("--ckpt_path", help="path to the checkpoint (TF2 prefix or directory)")
    ap.add_argument("--data_path", required=True, help="path to a .msgpack.zst frame file")
    ap.add_argument("--inflow", default=0, type=int, help="inflow timing")
    ap.add_argument("--timesteps", type=int, default=None)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--output_dir", default="./output")
    args, extra = ap.parse_known_args(argv)
    cfg = config.load_config(args.cfg_file, config.parse_cli_overrides(extra))
    model = config.build_model(cfg["model"])
    sim = Simulator(model, device=args.device)
    epoch = sim.load_ckpt(args.ckpt_path)
    data = datasets.load_msgpack_zst(args.data_path)
    steps = len(data) if args.timesteps is None else args.timesteps
    with torch.no_grad():
        results = run_rollout(sim, model, data[0], steps, args.inflow)
    pos = np.ones((len(results), results[-1].shape[0], 3)) * 1000  # like the reference: absent particles parked far away
    for i, r in enumerate(results):
        pos[i, :r.shape[0]] = r.cpu().numpy()
    out_dir = os.path.join(args.output_dir, "example", "0000")
    path = datasets.write_results(os.path.join(out_dir, "%04d.npz" % epoch), model.name,
                                  [(pos, {"name": "pred", "type": "PARTICLE"}),
                                   (np.asarray(data[0]["box"]), {"name": "bnd", "type": "PARTICLE"})])
    print(pos.shape, "->", path)
    return path


if __name__ == "__main__":
    sys.exit(0 if main() else 1)
