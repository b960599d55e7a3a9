python -c "
import cProfile, pstats, sys, io
sys.argv=['bench.py','--train-only','--train-dropout','0.2','--train-steps','6']
import runpy
pr=cProfile.Profile()
pr.enable()
try:
    runpy.run_path('bench.py', run_name='__main__')
except SystemExit:
    pass
pr.disable()
s=io.StringIO()
ps=pstats.Stats(pr,stream=s).sort_stats('tottime')
ps.print_stats(45)
open('gpurun_out/train_hostprof.txt','w').write(s.getvalue())
" > gpurun_out/train_hostprof.out 2>&1
