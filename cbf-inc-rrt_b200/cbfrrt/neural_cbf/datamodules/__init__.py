from .episodic_datamodule import EpisodicDataModule

__all__ = ["EpisodicDataModule"]
