class ConvergenceError(Exception):
    pass
