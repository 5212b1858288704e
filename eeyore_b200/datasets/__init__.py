from .data_counter import DataCounter
from .data_info import data_paths
from .xydataset import XYDataset
