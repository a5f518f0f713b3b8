"""bella_b200: B200-native (sm_100a) overlap-detection SpGEMM  C = A·Aᵀ  for the BELLA long-read
overlapper, behind a C-ABI (include/bella_b200.h).  See DESIGN.md."""
__all__ = ["frontend", "kmers", "pipeline", "spgemm", "xdrop"]
