# rounds per CUDA graph: is the 0.15 ms between the graph round (3.37 ms) and the sum of its kernels (3.22 ms) launch gaps?
timeout 600 python tools/explore_graph_rounds.py 16384 2>&1 | grep "rounds/graph\|probe"
