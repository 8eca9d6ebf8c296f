class ClusterData:
    """spmm_test.py:57-65 partitions AmazonProducts into <= 500 K-node parts with METIS; the stand-in yields the
    (already small) synthetic graph as its only parts."""

    def __init__(self, data, num_parts=1, save_dir=None, **_ignored):
        self.parts = [data, data]

    def __iter__(self):
        return iter(self.parts)
