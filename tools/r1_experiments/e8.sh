export SFC_FORCE_E=8
python - <<'PY'
import sys; sys.path.insert(0,'.')
import numpy as np, scipy.fft as sf
from scirs_b200 import FftPlan
from oracle import scirs2_fft_oracle as orc
rng=np.random.default_rng(0)
for (b,n,prec) in ((7,4096,'f64'),(5,2048,'f64'),(9,512,'f64'),(3,8192,'f64'),(7,4096,'f32')):
    x=(rng.standard_normal((b,n))+1j*rng.standard_normal((b,n))).astype(np.complex128 if prec=='f64' else np.complex64)
    y=FftPlan([b,n],[1],'c2c',prec,True).execute(x).reshape(b,n)
    print(n,prec,orc.rel_l2(y,sf.fft(x.astype(np.complex128),axis=1)))
x=rng.standard_normal((5,4096)); y=FftPlan([5,4096],[1],'r2c','f64').execute(x).reshape(5,2049); print('r2c',orc.rel_l2(y,sf.rfft(x,axis=1)))
z=FftPlan([5,4096],[1],'c2r','f64',scale=1/4096).execute(sf.rfft(x,axis=1)).reshape(5,4096); print('c2r',orc.rel_l2(z,x))
PY
python tools/gpu_bench.py c2c4096 rfft
echo "--- E=16"
unset SFC_FORCE_E
python tools/gpu_bench.py c2c4096
