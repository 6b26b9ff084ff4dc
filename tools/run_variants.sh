# builder tool (record of the end-of-round-2 occupancy experiment, profiles/README_r02.md item 11): the variant libraries
# are built first with tools/build_variant.py NAME flow.cu -D... into build/variants/ (not kept in the tree)
set -x
V=build/variants
python tools/front_time.py 524288
HGSFA_TC_CK=16 HGSFA_LIB=$V/libhgsfa_epi1ck16.so python tools/front_time.py 524288
HGSFA_TC_CK=16 HGSFA_TC_TIERS="128:56,256:113,512:227" HGSFA_LIB=$V/libhgsfa_epi0ck16b4.so python tools/front_time.py 524288
HGSFA_TC_CK=16 HGSFA_TC_TIERS="128:75,256:113,512:227" HGSFA_LIB=$V/libhgsfa_epi0ck16b4.so python tools/front_time.py 524288
HGSFA_TC_CK=16 HGSFA_TC_TIERS="128:75,256:113,512:227" HGSFA_LIB=$V/libhgsfa_epi1ck16b3.so python tools/front_time.py 524288
HGSFA_TC_CK=16 HGSFA_TC_TIERS="128:75,256:113,512:227" HGSFA_LIB=$V/libhgsfa_epi0ck16.so python tools/front_time.py 524288
