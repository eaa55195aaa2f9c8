"""Detector settings: defaults < detector.cfg < command line (thrifty/settings.py).

Only the settings the detect path consumes are defined (settings.py:23-109 keeps the same
keys, flags and defaults for them)."""
import logging
from collections import namedtuple

from thrifty_b200 import setting_parsers as sp

Definition = namedtuple("Definition", "args parser default description")

DEFINITIONS = {
    "sample_rate": Definition(["--sample-rate", "-s"], sp.metric_float, "2.4M", "Sample rate (sps)"),
    "chip_rate": Definition(["--chip-rate", "-p"], sp.metric_float, "0.999707M",
                            "Rate at which the code is being transmitted (bps)"),
    "tuner_freq": Definition(["--freq", "-f"], sp.metric_float, "433.83M", "Tuner center frequency (Hz)"),
    "tuner_gain": Definition(["--gain", "-g"], float, "0", "Tuner gain (dB)"),
    "capture_skip": Definition(["--skip", "-k"], int, "1", "Blocks to skip before capturing"),
    "block_size": Definition(["--block-size", "-b"], int, "16384",
                             "Length of fixed-sized blocks, a power of two (samples)"),
    "block_history": Definition(["--history", "-y"], int, "4920",
                                "Samples at the end of a block repeated at the start of the next"),
    "carrier_window": Definition(["--carrier-window", "-w"], sp.freq_range, "0--1",
                                 "Range of frequencies or frequency bins to look for carrier"),
    "carrier_threshold": Definition(["--carrier-threshold", "-t"], sp.threshold, "15*snr",
                                    "Threshold formula for carrier detector"),
    "corr_threshold": Definition(["--corr-threshold", "-u"], sp.threshold, "15*snr",
                                 "Threshold formula for correlation peak detector"),
    "template": Definition(["--template", "-z"], str, "template.npy", "Load template from a .npy file"),
    "rxid": Definition(["--rxid", "-r"], int, "-1", "Unique identifier of this receiver"),
}

DEFAULT_CONFIG_PATH = "detector.cfg"


class ConfigSyntaxError(Exception):
    def __init__(self, line_no, msg):
        Exception.__init__(self, "line #%d: %s" % (line_no, msg))


class SettingKeyError(KeyError):
    pass


class Namespace(dict):
    """dict with attribute access (settings.py:141-149)."""

    def __init__(self, d):
        dict.__init__(self, d)
        self.__dict__.update(d)


def parse_kvconfig(config_file):
    """'key: value' lines, '#' comments (settings.py:309-321)."""
    out = {}
    for line_no, line in enumerate(config_file):
        line = line.split("#", 1)[0]
        if not line.strip():
            continue
        if ":" not in line:
            raise ConfigSyntaxError(line_no + 1, "No delimiter found")
        key, value = line.split(":", 1)
        out[key.strip()] = value.strip()
    return out


def load(args=None, config_file=None, definitions=None):
    """Merge defaults, config file and argument strings, then parse (settings.py:170-231)."""
    definitions = definitions or DEFINITIONS
    strings = {k: d.default for k, d in definitions.items() if d.default is not None}
    for source in (parse_kvconfig(config_file) if config_file is not None else {}, args or {}):
        for key in source:
            if key not in definitions:
                raise SettingKeyError("Unknown setting: {}".format(key))
        strings.update(source)
    return {k: definitions[k].parser(v) for k, v in strings.items()}


def load_args(parser, keys, argv=None, definitions=None):
    """argparse front-end (settings.py:234-306): returns (settings Namespace, other-args Namespace)."""
    definitions = definitions or DEFINITIONS
    parser.add_argument("-v", "--verbose", action="store_true", help="Increase output verbosity")
    parser.add_argument("-c", "--config", dest="config", type=str, default=None,
                        help="Config file to load settings from [default: %s]" % DEFAULT_CONFIG_PATH)
    for key in keys:
        if key not in definitions:
            raise SettingKeyError("Unknown key: {}".format(key))
        d = definitions[key]
        parser.add_argument(*d.args, dest=key, type=str,
                            help="%s [default: %s]" % (d.description, d.default))
    args = vars(parser.parse_args(argv))
    if args["verbose"]:
        logging.basicConfig(level=logging.DEBUG)
    config_file = None
    if args["config"] is None:
        try:
            config_file = open(DEFAULT_CONFIG_PATH)
        except IOError:
            logging.warning("No config file found. Using default values.")
    else:
        config_file = open(args["config"])
    args.pop("config")
    key_args = {k: v for k, v in args.items() if k in keys and v is not None}
    extra = {k: v for k, v in args.items() if k not in keys}
    values = load(key_args, config_file, definitions)
    if config_file is not None:
        config_file.close()
    return Namespace({k: v for k, v in values.items() if k in keys}), Namespace(extra)
