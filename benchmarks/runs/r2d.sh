python benchmarks/variants.py --variants 0,1,0,1 --scene bench 2>&1 | cut -c1-400
