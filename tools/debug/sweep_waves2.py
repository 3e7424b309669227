import os, subprocess, sys
ROOT='/root/repo'
for fpw, lanes, iters in [(64,3,None),(64,4,16),(64,4,12),(64,4,24),(64,5,16),(64,6,16),(64,5,24),(48,4,16),(96,4,16),(64,3,None),(64,4,16)]:
    env=dict(os.environ)
    if iters is not None: env["JPEGB200_TK_ITERS"]=str(iters)
    r=subprocess.run([sys.executable, os.path.join(ROOT,'tools/debug/sweep_waves.py'),'--child','--batch','1024','--fpw',str(fpw),'--lanes',str(lanes)],env=env,capture_output=True,text=True,timeout=600)
    line=[l for l in r.stdout.splitlines() if l.startswith('{')]
    print(line[-1] if line else 'FAILED '+r.stderr[-200:], flush=True)
