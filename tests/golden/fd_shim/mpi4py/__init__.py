class MPI:
    MIN = "min"
    MAX = "max"
