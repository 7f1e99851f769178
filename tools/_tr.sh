python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/dev_chain.py --check 1 --iters 10 | grep "A^" | awk '{print $1,$4,$5,$6,$NF}' | tr '\n' ' '; echo
B200_NARROW=0 python tools/dev_chain.py --check 0 --iters 10 | grep "A^" | awk '{print $1,$4,$5,$6}' | tr '\n' ' '; echo
python tools/big_config.py torus --side 100 --power 5 --check 1 --iters 3 2>&1| grep -o "^A^[0-9]\|ms=[0-9.]*\|BIT-EXACT\|MISMATCH" | paste - - - | tr '\n' ' '; echo
python tools/big_config.py rmat --scale 18 --check 1 --iters 3 2>&1| grep -o "^A^[0-9]\|ms=[0-9.]*\|BIT-EXACT\|MISMATCH\|mode=[0-9]" | paste - - - - | tr '\n' ' '; echo
