"""Write a synthetic `.card` file with the SURVEY.md 8d block distribution."""
import argparse

import numpy as np

from thrifty_b200 import block_data, synth


def _main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("output", type=argparse.FileType("w"))
    ap.add_argument("--template", required=True, help=".npy template")
    ap.add_argument("--blocks", type=int, default=1000)
    ap.add_argument("--block-size", type=int, default=16384)
    ap.add_argument("--history", type=int, default=4920)
    ap.add_argument("--p-signal", type=float, default=0.5)
    ap.add_argument("--seed", type=int, default=synth.SEED0)
    args = ap.parse_args(argv)
    tpl = np.load(args.template)
    raw, _ = synth.make_blocks(args.blocks, args.block_size, args.history, tpl, args.p_signal, seed=args.seed)
    block_data.write_card(args.output, raw, header={"history_size": args.history})
    args.output.close()


if __name__ == "__main__":
    _main()
