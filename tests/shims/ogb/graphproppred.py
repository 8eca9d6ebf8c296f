class PygGraphPropPredDataset:
    def __init__(self, *a, **k):
        raise NotImplementedError("graph-property datasets are not on the aggregation path")
