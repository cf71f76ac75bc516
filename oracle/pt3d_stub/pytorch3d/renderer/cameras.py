class CamerasBase:  # type annotation only on the in-tree path
    pass
