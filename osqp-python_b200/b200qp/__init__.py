"""b200qp -- helpers around the engine: synthetic problems and a ctypes view of the C ABI."""
