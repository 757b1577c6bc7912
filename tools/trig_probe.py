import sys; sys.path.insert(0,'.')
from botlab_b200 import engine
e = engine.Engine(1024)
for lo,hi in [(-3.2,3.2),(-9.5,9.5),(-9.5,3.2),(-6.3,6.3)]:
    print(lo,hi,e.fast_trig_error(lo,hi))
