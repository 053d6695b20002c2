echo "== TS (A in TMEM, 2 CTAs/SM planned)"; timeout 300 python scratch/f_experiments.py 2>&1 | tail -9
echo "== SS (A in smem, 1 CTA/SM)"; CIRS_F_ATM=0 timeout 300 python scratch/f_experiments.py 2>&1 | tail -9
