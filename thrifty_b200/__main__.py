from thrifty_b200.cli import _main

_main()
