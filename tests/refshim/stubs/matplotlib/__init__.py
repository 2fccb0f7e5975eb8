"""Import stub (TEST INFRASTRUCTURE): deluca/envs/_lds.py imports matplotlib.pyplot for a plot helper that is never called."""
