import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'vector-indexer_b200'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np
import oracle as O
from vector_indexer_py import _ffi as ffi
from conftest import bench_data
from test_gpu_search import make_pair
d=int(sys.argv[1]); k=int(sys.argv[2]); npb=int(sys.argv[3]); n=int(sys.argv[4]) if len(sys.argv)>4 else 30000; nl=int(sys.argv[5]) if len(sys.argv)>5 else 12
xb,xq=bench_data(n,d,3000)
oix,gix=make_pair(O,ffi,xb,nl)
try:
    Dg,Ig=gix.search(xq,k,npb)
    Do,Io=oix.search_batch(xq,k,npb,nthreads=0)
    print(d,k,npb,n,nl,'ok', np.array_equal(Dg.view(np.uint32),Do.view(np.uint32)), np.array_equal(Ig,Io))
except Exception as e:
    print(d,k,npb,n,nl,'FAIL',str(e)[:100])
