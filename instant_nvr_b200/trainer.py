"""Trainer shim: ``trainer_module: instant_nvr_b200.trainer`` (``lib/train/trainers/make_trainer.py:4-12``).

The reference's ``NetworkWrapper`` hard-codes its own renderer
(``lib/train/trainers/inb_trainer.py:23``: ``self.renderer = inb_renderer.Renderer(self.net)``), so pointing
``renderer_module`` at this package is not enough for training.  This wrapper is the reference's, with the renderer
replaced by the B200 one; every loss term, statistic and optimizer interaction stays the reference's code.

Importable only inside the reference tree (it subclasses ``lib.train.trainers.inb_trainer.NetworkWrapper``).
"""
from lib.train.trainers import inb_trainer as _ref   # noqa: E402  (reference package)

from .renderer import Renderer


class NetworkWrapper(_ref.NetworkWrapper):
    def __init__(self, net):
        super().__init__(net)
        # training outputs stay on the device and attached to autograd; eval outputs go to the CPU as the
        # reference's renderer does (inb_renderer.py:199-200)
        self.renderer = Renderer(self.net)
