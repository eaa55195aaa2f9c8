"""Test infrastructure only: CPU restatements of the reference algorithm used as the parity checker (see oracle/README or the module headers).  Nothing under thrifty_b200/ imports this package."""
