"""Command line of reference main/hive.py for the headless modes (blind run, benchmark).

    python -m game_of_flies_b200.hive -b -r output/sim_dir      # 1000 steps, record frames
    python -m game_of_flies_b200.hive -be -nt 1000 1000         # 500 steps of a 2D Clusters run

The interactive mode (-i) is the reference's PyQt/VisPy UI; it is not rebuilt here.  Point
its `simulation.py` at game_of_flies_b200.simulation to use this engine.
"""
from __future__ import annotations

import sys
from argparse import ArgumentParser

import numpy as np


def parse(argv=None):
    p = ArgumentParser()
    p.add_argument("-b", "--blind", action="store_true", default=False, help="Run simulation without visualisation")
    p.add_argument("-i", "--ui", action="store_true", default=False, help="Run simulation with UI (reference only)")
    p.add_argument("-r", "--record", type=str, default=None, help="Path to directory to store simulation")
    p.add_argument("-be", "--benchmark", action="store_true", default=False, help="For Benchmark")
    p.add_argument("-nt", "--ntype", type=int, nargs="+", default=[100, 100], help="num of particles")
    # the reference also accepts `key=value` spellings (hive.py:41-49)
    args = [tok for a in (sys.argv[1:] if argv is None else argv) for tok in a.replace("=", " ").split()]
    return p.parse_args(args)


def main(argv=None):
    from .simulation import Simulation
    args = parse(argv)
    if args.ui or not (args.blind or args.benchmark):
        raise SystemExit("the PyQt/VisPy UI lives in the reference; see INTEGRATION.md to run it on this engine")
    if args.benchmark:
        sim = Simulation(np.array(args.ntype), np.array([0]), limits=(100, 100, 0), seed=434)
        sim.update()
        sim.bench_run(500, args.record)
    else:
        sim = Simulation(np.array([2000] * 4), np.array([2000] * 3), limits=(100, 100, 100), seed=434)
        sim.update()
        sim.blind_run(1000, args.record)


if __name__ == "__main__":
    main()
