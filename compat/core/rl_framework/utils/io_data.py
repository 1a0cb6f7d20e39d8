"""The reference's module path src/core/rl_framework/utils/io_data.py, backed by dronechase_b200.io_data."""
from dronechase_b200.io_data import DatasetWriter, IOData, MultiFileDataset, collect_data, collect_data_multiobs  # noqa: F401

MultiH5Dataset = MultiFileDataset      # io_data.py:13-52
