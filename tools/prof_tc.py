import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.debug_tc import run
run(2048, float(os.environ.get("THETA", "6.0")), nscreens=int(os.environ.get("NS", "8")))
