from .virtual_radar import VirtualRadar, edges  # noqa: F401
