"""pyspark.ml.evaluation.MulticlassClassificationEvaluator (kdd99.py:86-91; cicids17.py:90-95):
confusion counts by the b200flow kernel (R10), metrics per MulticlassMetrics (A.8) + macro-F1."""
import torch

from b200flow import dist as bdist
from b200flow import forest as fr

from .param import Params


class MulticlassClassificationEvaluator(Params):
    _defaults = {"predictionCol": "prediction", "labelCol": "label", "metricName": "f1"}
    _metrics = ("f1", "accuracy", "weightedPrecision", "weightedRecall", "macroF1")

    def __init__(self, predictionCol=None, labelCol=None, metricName=None):
        super().__init__(predictionCol=predictionCol, labelCol=labelCol, metricName=metricName)

    def confusionMatrix(self, dataset):
        pred = dataset._column_tensor(self.getOrDefault("predictionCol")).to(torch.float64).contiguous()
        lab = dataset._column_tensor(self.getOrDefault("labelCol")).to(torch.float64).contiguous()
        # an empty local shard still takes part in both collectives (every rank issues the same sequence)
        mx = (torch.maximum(pred.max(), lab.max()) if pred.numel() else torch.zeros((), dtype=torch.float64, device=pred.device)).reshape(1)
        if bdist.group() is not None:
            import torch.distributed as dist
            bdist.all_reduce_(mx, bdist.group(), op=dist.ReduceOp.MAX)
        C = int(mx.item()) + 1
        return bdist.all_reduce_sum_(fr.confusion_matrix(pred, lab, C)).cpu()

    def evaluate(self, dataset, params=None):
        ev = self.copy(params) if params else self
        name = ev.getOrDefault("metricName")
        if name not in self._metrics:
            raise ValueError("metricName must be one of %s, got %r" % (list(self._metrics), name))
        return fr.metrics_from_confusion(ev.confusionMatrix(dataset).numpy())[name]

    def isLargerBetter(self):
        return True
