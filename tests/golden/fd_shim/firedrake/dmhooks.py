def get_function_space(*a, **k):
    raise NotImplementedError


def get_appctx(*a, **k):
    raise NotImplementedError


def push_appctx(*a, **k):
    raise NotImplementedError
