"""B200-native structure-branch hot path of Metagenomic-DeepFRI.

Drop-in surface (same names / argument meaning as the reference):
  contact_map_utils.pairwise_sqeuclidean / align_contact_map   (mDeepFRI/contact_map_utils.pyx:17,44)
  bio_utils.calculate_contact_map / build_align_contact_map    (mDeepFRI/bio_utils.py:196,348)
  contact_map.CAlphaCoordinates / DistanceMap / ContactMap     (mDeepFRI/contact_map.py)
  predict.seq2onehot / Predictor                               (mDeepFRI/predict.pyx:17,50)

Everything computes on an sm_100a GPU through the C-ABI library `libmdf_b200.so`
(`include/mdf_b200.h`).  There is no CPU fallback: a missing library or GPU raises.

The directory name carries a hyphen, so import it through `mdf_pkg.load()` at the repo
root (registers it as `metagenomic_deepfri_b200`).
"""
__version__ = "0.1.0"
