for th in 8 16 32; do timeout 200 python -c "
import sys; sys.path.insert(0,'.')
import bench, time
t=time.time(); v,c,ms=bench.cpu_reference(1,0,threads=$th); print('threads',$th,'pairs/s',round(v,4),'ms/step',round(ms), flush=True)" 2>/dev/null; done
