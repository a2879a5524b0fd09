class PETSc:
    pass
