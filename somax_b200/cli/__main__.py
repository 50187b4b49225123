"""`python -m somax_b200.cli ...` = the `somax-sim` command line."""
import sys

from .app import main

sys.exit(main())
