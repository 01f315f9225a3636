from .build import build_offset_regression
