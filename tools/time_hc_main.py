import os, subprocess, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trifocal_pose_estimation_using_improved_gpuhc_b200 import fixtures
exe=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),'trifocal_pose_estimation_using_improved_gpuhc_b200','lib','hc-main')
with tempfile.TemporaryDirectory() as root:
    fixtures.materialize_tree(root, files=[0])
    for H in (100, 10000, 100000):
        t=time.time()
        out=subprocess.run([exe,'-p','trifocal_2op1p_30x30','-s','Num_Of_RANSAC_Iterations=%d'%H,'-s','Verbose=true'],cwd=os.path.join(root,'build','bin'),capture_output=True,text=True,timeout=900)
        dt=time.time()-t
        lines=[l for l in out.stdout.splitlines() if 'GPU Computation Time =' in l or 'Phases' in l or 'Driver wall' in l]
        print(H, 'rc',out.returncode,'wall %.2f s'%dt, '|', ' | '.join(l.strip() for l in lines[:8]))
