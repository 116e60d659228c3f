for mb in 16 32 64 128 256; do for sl in 2 3 4; do
VFGS_B200_CHUNK_MB=$mb VFGS_B200_SLOTS=$sl python bench.py --no-cpu-baseline --steps 3 --warmup 3 --frames-per-step 32 --e2e-frames 32 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk_mb',$mb,'slots',$sl,'e2e', round(d['e2e']['value']), 'fps', round(d['e2e']['value']*49766400*2/2/1e9,1),'GB/s per dir')"
done; done
