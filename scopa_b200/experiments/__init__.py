"""Experiment runners that emit the reference's result-file shapes."""
