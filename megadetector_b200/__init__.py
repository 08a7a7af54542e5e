"""
megadetector_b200: a B200-native (sm_100a) batch detection engine that drops in behind MegaDetector's
`PTDetector` / `run_detector.load_detector` / `run_detector_batch`.  See DESIGN.md and INTEGRATION.md.
"""

__version__ = '0.1.0'


def load_detector(model_path, detector_options=None, verbose=False):
    """Counterpart of run_detector.load_detector() for `.pt` YOLOv5-P6 models (run_detector.py:601-681)."""
    from .detector import B200Detector
    return B200Detector(model_path, detector_options, verbose=verbose)
