"""Accuracy of mcmcb_exp_fast on the GPU vs mpmath-free reference (numpy exp is ~0.5 ulp)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
