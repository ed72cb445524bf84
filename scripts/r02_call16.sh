set -x
CO=$PWD/asvd4llm_b200/csrc/libasvd_b200_co.so
timeout 120 python scripts/ab_co.py 4096x4096 18
ASVD_B200_LIBPATH=$CO ASVD_B200_OVERLAP=1 ASVD_B200_SOLVE=lean timeout 120 python scripts/ab_co.py 4096x4096 18
ASVD_B200_LIBPATH=$CO ASVD_B200_OVERLAP=1 ASVD_B200_SOLVE=lean timeout 120 python scripts/ab_co.py 4096x4096 32
ASVD_B200_LIBPATH=$CO ASVD_B200_OVERLAP=1 ASVD_B200_SOLVE=lean timeout 120 python scripts/ab_co.py 11008x4096 18
timeout 120 python scripts/ab_co.py 11008x4096 18
