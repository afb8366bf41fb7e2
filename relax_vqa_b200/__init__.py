"""relax_vqa_b200 - B200-native (sm_100a) implementation of the ReLaX-VQA feature-extraction
hot path behind the reference's Python entry points.  See DESIGN.md."""
__version__ = "0.1.0"
