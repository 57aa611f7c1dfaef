"""Result / message channel of the iLQR path.

Mirrors the surface of the reference's singleton logger
(/root/reference/CuriousRL/utils/Logger.py:8-124): numeric level gate
(`set_level`), `info/debug/warn/error`, and the `save_to_json` /
`read_from_json` pair that `scenario.play()` uses to fetch the last
trajectory (Logger.py:66-99).  Built on the stdlib `logging` module; file
sinks are created lazily, and only when `set_is_save_json(True)` or a log
folder is asked for.

Difference kept on purpose: `save_to_json` does not serialise when nothing is
saved to disk (the reference runs `json.dumps` on every call, Logger.py:69,
which costs as much as a whole numeric iteration - SURVEY.md section 5).
"""
from __future__ import annotations

import json
import logging
import os
import shutil
import sys
from datetime import datetime


class Logger:
    DEBUG = 10
    INFO = 20
    WARN = 30
    ERROR = 40
    DISABLED = 50

    _instance = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
            cls._instance._setup()
        return cls._instance

    def _setup(self):
        self.MIN_LEVEL = 0
        self.IS_USE_LOGGER = True
        self.IS_SAVE_JSON = False
        self.FOLDER_NAME = datetime.now().strftime("%Y_%m_%d_%H_%M_%S_%f")
        self.memory_json = None
        self.memory_batch = None
        self._py = logging.getLogger("curiousrl_b200")
        self._py.setLevel(logging.DEBUG)
        self._py.propagate = False
        if not self._py.handlers:
            h = logging.StreamHandler(sys.stdout)
            h.setFormatter(logging.Formatter("%(asctime)s - %(levelname)s: %(message)s"))
            self._py.addHandler(h)
        self._file_handler = None

    # ------------------------------------------------------------------ sinks
    @property
    def logger_path(self):
        if not self.IS_USE_LOGGER:
            return None
        return os.path.join("logs", self.FOLDER_NAME)

    def _ensure_file_sink(self):
        if self._file_handler is None and self.IS_USE_LOGGER:
            os.makedirs(self.logger_path, exist_ok=True)
            self._file_handler = logging.FileHandler(os.path.join(self.logger_path, "main.log"), delay=True)
            self._file_handler.setFormatter(logging.Formatter("%(asctime)s - %(levelname)s: %(message)s"))
            self._py.addHandler(self._file_handler)

    def _drop_file_sink(self):
        if self._file_handler is not None:
            self._py.removeHandler(self._file_handler)
            self._file_handler.close()
            self._file_handler = None

    def set_is_use_logger(self, is_use_logger):
        self.IS_USE_LOGGER = bool(is_use_logger)
        if not self.IS_USE_LOGGER:
            self._drop_file_sink()
        return self

    def set_folder_name(self, folder_name, remove_existing_folder=True):
        self._drop_file_sink()
        path = os.path.join("logs", folder_name)
        if os.path.isdir(path):
            if remove_existing_folder:
                shutil.rmtree(path)
                self.info('[+] The folder "%s" existed and was removed.' % folder_name)
            else:
                self.info('[+] The folder "%s" exists; log lines are appended.' % folder_name)
        self.FOLDER_NAME = folder_name
        return self

    def set_is_save_json(self, is_save_json):
        self.IS_SAVE_JSON = bool(is_save_json)
        return self

    # ------------------------------------------------------- result channel
    def save_to_json(self, **kwargs):
        """Keep the record in memory, or append one JSON line to logs/<folder>/main.json."""
        if self.IS_SAVE_JSON:
            if not self.IS_USE_LOGGER:
                raise Exception('Logger must be used to save json. Please call "logger.set_is_use_logger(True)".')
            os.makedirs(self.logger_path, exist_ok=True)
            with open(os.path.join(self.logger_path, "main.json"), "a") as fh:
                fh.write(json.dumps(kwargs) + "\n")
        else:
            self.memory_json = kwargs

    def read_from_json(self, folder_name=None, no_iter=-1):
        if not self.IS_SAVE_JSON and folder_name is None:
            return self.memory_json
        if folder_name is None:
            folder_name = self.FOLDER_NAME
        path = os.path.join("logs", folder_name, "main.json")
        if not os.path.isfile(path):
            raise Exception('The json file is not saved in the log folder "%s". '
                            'Call logger.set_is_save_json(True) before solve().' % folder_name)
        last = None
        with open(path) as fh:
            for i, line in enumerate(fh):
                last = line
                if i == no_iter:
                    return json.loads(line)
        if no_iter == -1 and last is not None:
            return json.loads(last)
        raise Exception("The number of iteration exceeds the maximum iteration number!")

    # ------------------------------------------- batched result channel
    def save_batch(self, meta=None, **arrays):
        """Batched counterpart of `save_to_json` (utils/result_stream.py): named arrays with a leading batch axis.
        In memory (default) the record is kept by reference - device tensors stay on the device; with
        `set_is_save_json(True)` it is appended to the binary stream under logs/<folder>/stream/."""
        if self.IS_SAVE_JSON:
            if not self.IS_USE_LOGGER:
                raise Exception('Logger must be used to save json. Please call "logger.set_is_use_logger(True)".')
            from .result_stream import ResultStream
            return ResultStream(self.FOLDER_NAME).append(meta, **arrays)
        self.memory_batch = dict(arrays, meta=meta or {})
        return None

    def read_batch(self, folder_name=None, no_iter=-1):
        """Batched counterpart of `read_from_json`: record `no_iter` (-1 = last) as {name: array}."""
        if not self.IS_SAVE_JSON and folder_name is None:
            return self.memory_batch
        from .result_stream import ResultStream
        return ResultStream(self.FOLDER_NAME if folder_name is None else folder_name).read(no_iter)

    # ------------------------------------------------------------- messages
    def set_level(self, level):
        self.MIN_LEVEL = level

    def _emit(self, level, py_level, args):
        if self.MIN_LEVEL <= level:
            if self.IS_USE_LOGGER and self.IS_SAVE_JSON:
                self._ensure_file_sink()
            self._py.log(py_level, *args)

    def debug(self, *args):
        self._emit(self.DEBUG, logging.DEBUG, args)

    def info(self, *args):
        self._emit(self.INFO, logging.INFO, args)

    def warn(self, *args):
        self._emit(self.WARN, logging.WARNING, args)

    def error(self, *args):
        self._emit(self.ERROR, logging.ERROR, args)


logger = Logger()
