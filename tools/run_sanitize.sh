timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize.py 2>&1 | tail -25
echo "memcheck rc=$?"
MPPIB_STREAM=1 timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "
import sys; sys.path.insert(0,'.')
import mppi_generic_b200 as m
from mppi_generic_b200 import workloads as W
for name in ('racer_lstm','cartpole'):
    w=W.by_name(name,256,64); e=w.make_engine(); e.solve(w.x0,w.U0); e.solve(w.x0,w.U0); e.close(); print('race ok',name)
" 2>&1 | tail -12
