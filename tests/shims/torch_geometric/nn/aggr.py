import torch


class Aggregation(torch.nn.Module):
    def reset_parameters(self):
        pass


class SumAggregation(Aggregation):
    pass


class MultiAggregation(Aggregation):
    def get_out_channels(self, in_channels):
        return in_channels
