class Citations:
    pass
