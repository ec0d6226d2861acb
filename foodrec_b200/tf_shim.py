"""The handful of ``tf.*`` names ``Train_recommender.py`` touches outside the model
(``:117-121,145-151,223``), so the reference driver runs with
``import foodrec_b200.tf_shim as tf`` and ``from foodrec_b200 import Model, evaluate_model``
(see INTEGRATION.md).  Nothing here computes anything."""
from .model import ConfigProto, Saver, Session, global_variables_initializer, latest_checkpoint


class train:  # noqa: N801  (mirrors tf.train)
    Saver = Saver
    latest_checkpoint = staticmethod(latest_checkpoint)


__all__ = ["ConfigProto", "Session", "global_variables_initializer", "train"]
