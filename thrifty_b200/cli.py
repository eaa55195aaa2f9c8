"""`python -m thrifty_b200 <command>`: sub-command dispatch in the style of thrifty/cli.py:47-92.

Only the detect path is provided (plus a synthetic `.card` generator for tests/benchmarks)."""
import importlib
import sys

HELP = """usage: thrifty_b200 <command> [<args>]

    detect   Detect positioning signals in .card / raw data and estimate SoA (GPU)
    fastdet  Same job with the semantics and options of the reference's native `fastdet` (GPU)
    identify Merge .toad files, identify transmitters by carrier bin, drop duplicates -> .toads (CUDA kernels)
    synth    Write a synthetic .card file (see SURVEY.md 8d)

Use 'thrifty_b200 help <command>' for a command's arguments."""

MODULES = {"detect": "thrifty_b200.detect", "fastdet": "thrifty_b200.fastdet", "identify": "thrifty_b200.identify",
           "synth": "thrifty_b200.synth_cli"}


def _main():
    if len(sys.argv) == 1:
        print(HELP)
        sys.exit(1)
    command = sys.argv.pop(1)
    if command in ("help", "--help"):
        if len(sys.argv) == 2:
            command = sys.argv.pop(1)
            sys.argv.append("--help")
        else:
            print(HELP)
            sys.exit(0)
    if command not in MODULES:
        print("thrifty_b200: {} is not a command. See 'thrifty_b200 --help'.".format(command),
              file=sys.stderr)
        sys.exit(1)
    sys.argv[0] += " " + command
    importlib.import_module(MODULES[command])._main()


if __name__ == "__main__":
    _main()
