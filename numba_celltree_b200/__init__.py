"""B200-native drop-in for the numba_celltree query hot path (CellTree2d / EdgeCellTree2d)."""

__version__ = "0.1.0"

__all__ = ("CellTree2d", "EdgeCellTree2d", "trim_memory")


def trim_memory() -> None:
    """Give the library's cached device blocks and page-locked host blocks back to the driver."""
    from numba_celltree_b200 import _lib

    lib = _lib.load()
    lib.ct_device_trim()
    lib.ct_host_trim()


def __getattr__(name):
    # Lazy: importing the package (e.g. for numba_celltree_b200.synthetic) must not need the CUDA library.
    if name == "CellTree2d":
        from numba_celltree_b200.celltree import CellTree2d

        return CellTree2d
    if name == "EdgeCellTree2d":
        from numba_celltree_b200.edge_celltree import EdgeCellTree2d

        return EdgeCellTree2d
    raise AttributeError(name)
