"""``kdsource-resample KDEFILE OUTMCPL NPARTICLES [-s SEED] [-f]`` (reference
``python/kdsource/_cli_resample.py``)."""

import argparse


def parse_args(argv=None):
    p = argparse.ArgumentParser(
        description="Generate new MCPL files based upon an existing kernel density "
                    "estimate XML file.")
    p.add_argument("kdefile", metavar="KDEFILE",
                   help="Kernel Density Estimate .xml file generated via the Python interface.")
    p.add_argument("outmcpl", metavar="OUTMCPL", help="Name of MCPL file which will be created.")
    p.add_argument("nparticles", metavar="NPARTICLES", type=int,
                   help="Number of particles to sample and place in OUTMCPL.")
    p.add_argument("-s", "--seed", metavar="SEED", type=int,
                   help="Seed of the random stream (default: from os.urandom).")
    p.add_argument("-f", "--force", action="store_true",
                   help="Will overwrite existing file if it already exists.")
    return p.parse_args(argv)


def main(argv=None):
    from .resample import resample_to_mcpl

    args = parse_args(argv)
    resample_to_mcpl(args.kdefile, args.outmcpl, args.nparticles, force=args.force,
                     rng=args.seed)


if __name__ == "__main__":
    main()
