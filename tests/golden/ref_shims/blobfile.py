def BlobFile(path, mode="rb"):
    return open(path, mode)
def exists(path):
    import os
    return os.path.exists(path)
