"""Makes `python examples/<name>.py` work from a source checkout (puts the repository root on sys.path)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
