"""Entry script: run the Causal Bayesian Optimisation agent (same three imports and four steps as the reference's
runCBO.py, which also runs unchanged against this repository's `src` package)."""
from src.CBO import *  # noqa: F401,F403
from src.ArgumentParser import ArgumentParser as ArgumentParser
from src.DataLoader import DataLoader as DataLoader


def main():
    args = ArgumentParser().parse(verbose=True)                          # flags + numpy seed
    data = DataLoader(args.experiment, args.initial_num_obs_samples)     # measurements, interventions, graph
    CBO(args, data, verbose=True).run()                                  # observe / intervene loop


if __name__ == "__main__":
    main()
