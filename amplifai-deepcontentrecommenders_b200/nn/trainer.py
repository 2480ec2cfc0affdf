"""Abstract trainer interface (mirrors dcrecommend/nn/trainer.py)."""
from abc import ABC, abstractmethod


class Trainer(ABC):

    """Interface every trainer implements: fit / predict / score / save."""

    @abstractmethod
    def fit(self, *args, **kwargs):
        """Train the model."""

    @abstractmethod
    def predict(self, *args, **kwargs):
        """Score candidates for one entity."""

    @abstractmethod
    def score(self, *args, **kwargs):
        """Evaluate the model."""

    @abstractmethod
    def save(self, *args, **kwargs):
        """Persist the trainer state."""
