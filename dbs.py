#!/usr/bin/env python
"""Reference-compatible entry point: ``python dbs.py -d false -ws 4 -b 512 -m densenet -ds cifar10 -gpu 0,1,2,3``
(same flags as the reference's ``dbs.py``; see ``dynamic_load_balance_distributeddnn_b200/cli.py``)."""
import sys

from dynamic_load_balance_distributeddnn_b200.cli import main

if __name__ == "__main__":
    sys.exit(main())
