from .auroc import BinaryAUROCMetric
from .confusion_matrix import ConfusionMatrixMetric, ConfusionMatrixMetricBuilder, confusion_matrix_metric

__all__ = ["BinaryAUROCMetric", "ConfusionMatrixMetric", "ConfusionMatrixMetricBuilder", "confusion_matrix_metric"]
