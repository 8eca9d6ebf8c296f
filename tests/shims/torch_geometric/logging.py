def init_wandb(*a, **k):
    pass


def log(**kwargs):
    print(", ".join("%s: %s" % kv for kv in kwargs.items()))
