python tests/mgpu_worker.py --out /tmp/w1.npz --particles 40001 --steps 3
python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29552 tests/mgpu_worker.py --out /tmp/w2.npz --particles 40001 --steps 3 2>&1 | tail -3
python - <<PY
import numpy as np
a=np.load('/tmp/w1.npz'); b=np.load('/tmp/w2.npz')
ca, cb = a['cloud'], b['cloud']
for k in ('pose','parent_pose'):
    for f in ('utime','x','y','theta'):
        d = np.nonzero(ca[k][f] != cb[k][f])[0]
        print(k, f, 'ndiff', len(d), d[:8], ca[k][f][d[:4]], cb[k][f][d[:4]])
d=np.nonzero(ca['weight']!=cb['weight'])[0]; print('weight ndiff', len(d), d[:8], ca['weight'][d[:3]], cb['weight'][d[:3]])
print(a['estimates'], b['estimates'])
PY
