import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.microbench import run
run(3, 2, 60, 220, 85, reps=30)
