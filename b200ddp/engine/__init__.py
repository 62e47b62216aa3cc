from .step import TrainStep
from .trainer import Trainer, build_dataset, build_criterion
from .cli import build_parser, setup, cleanup, main

__all__ = ["TrainStep", "Trainer", "build_dataset", "build_criterion", "build_parser", "setup", "cleanup", "main"]
