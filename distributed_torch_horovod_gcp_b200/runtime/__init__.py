"""C++/CUDA runtime: symmetric-memory heap over cuMem VMM + NVLS multicast, fd exchange,
signal pads, and the ctypes loader for the in-tree shared libraries."""
