"""re-export of oracle/pyoracle.py for the tests"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.pyoracle import *  # noqa: F401,F403,E402
from oracle.pyoracle import lib, ROOT  # noqa: F401,E402
