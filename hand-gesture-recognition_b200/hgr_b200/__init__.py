"""hgr_b200 - B200-native forward path of yingkunwu/hand-gesture-recognition's MultiTaskNet."""
from .model import MultiTaskNet  # noqa: F401
from .ops import crop_normalize, crop_warp_normalize, get_max_preds, pose_accuracy  # noqa: F401
from .pipeline import HandPipeline, ShardedHandPipeline  # noqa: F401
from .serving import ClassifierSession  # noqa: F401
from .training import DataParallelTrainer, loss_and_grads  # noqa: F401

__all__ = ["MultiTaskNet", "HandPipeline", "ShardedHandPipeline", "DataParallelTrainer", "ClassifierSession", "loss_and_grads", "get_max_preds", "crop_normalize", "crop_warp_normalize", "pose_accuracy"]
