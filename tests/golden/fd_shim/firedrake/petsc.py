class PETSc:  # placeholder: preconditioners.py is imported but never executed by the shim
    pass
