def map(f, tree):  # noqa: A001
    if isinstance(tree, (tuple, list)):
        return type(tree)(map(f, t) for t in tree)
    return f(tree)
