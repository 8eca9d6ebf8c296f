class PygLinkPropPredDataset:
    def __init__(self, *a, **k):
        raise NotImplementedError("link-property datasets are not on the aggregation path")
