python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/dev_chain.py --check 1 --iters 8 | grep "A^" | awk '{print $1,$6,$NF}' | tr '\n' ' '; echo
B200_HOSTTIME=1 python tools/dev_chain.py --check 0 --iters 3 --steps 3 2>&1 | grep -E "hosttime" | tail -2 | cut -c1-260
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-330
