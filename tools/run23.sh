cd $GRAFT_REPO_ROOT
python tools/idx_probe.py dense
python tools/idx_probe.py nodense
python tools/idx_probe.py sparse
PW_BUCKET_IDXMUL=8 python tools/idx_probe.py sparse
PW_BUCKET_J=10 PW_BUCKET_STAGES=0 python tools/idx_probe.py sparse
