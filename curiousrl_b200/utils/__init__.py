from .Logger import logger, Logger
