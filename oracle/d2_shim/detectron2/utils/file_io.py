"""detectron2.utils.file_io.PathManager for local files (iopath's PathManager restated for the two calls the reference's
evaluator makes: coin/evaluation/cloud_pascal_voc_evaluation.py:37,147,230)."""


class _LocalPathManager:
    @staticmethod
    def open(path, mode="r", **kw):
        return open(path, mode, **kw)

    @staticmethod
    def get_local_path(path, **kw):
        return path


PathManager = _LocalPathManager()
