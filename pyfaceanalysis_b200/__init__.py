"""pyfaceanalysis_b200 -- B200-native (sm_100a) implementation of the HiGSFA sliding-window hot path of
AlbertoEsc/PyFaceAnalysis: window extraction -> flow forward -> Gaussian-classifier heads.

Public surface (mirrors the reference's objects):

    GpuFlow(flow_obj).execute(x, benchmark=None)          # networks[i].execute
    GpuGaussianClassifier(clf_obj).regression(x, labels)  # classifiers[i].regression / .label
    load_network_subimages(...) / extract_subimages(...)  # face_analysis.load_network_subimages
    load_obj(base_dir, base_filename)                     # Cache.load_obj_from_cache
    AttributeEstimator(net, age, race, gender).estimate   # estimate_age_race_gender on normalised crops
    cascade.FaceDetector(...).detect(images)              # the per-image loop of FaceDetectUpdated.py, batched
    batch.run_batch(detector, batch_filename)             # python FaceDetect.py --batch=batch_filename

All compute goes through ``libhgsfa.so`` (``include/hgsfa.h``); there is no CPU fallback.
"""
from .pickles import load_obj, loads as unpickle  # noqa: F401
from .gpuflow import GpuFlow, GpuNode  # noqa: F401
from .classifier import GpuGaussianClassifier  # noqa: F401
from .crop import extract_subimages, load_network_subimages, NEAREST, BILINEAR, BICUBIC  # noqa: F401
from .attributes import AttributeEstimator  # noqa: F401

__version__ = "0.1.0"
