"""`python -m nanocaller_b200 ...` = the NanoCaller command line over the B200 path (nanocaller_b200/cli.py)."""
from .cli import main

main()
