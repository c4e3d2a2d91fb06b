"""The ONE file a RUMpy maintainer adds to adopt rumpy_b200: dropped at
`rumpy/SISR/models/b200/handlers.py` inside the RUMpy tree, it is found by the registry's AST scan
(rumpy/shared_framework/models/__init__.py:7-25), which maps every `class XHandler` to the model name `x`.
Nothing else in RUMpy changes: `[model] name = "rcanb200"` in a TOML config (or `new_params={'name': 'rcanb200', ...}`
for `SISRInterface`) selects the sm_100a path behind the reference's own interface, trainer and evaluation hub."""
from rumpy_b200.SISR.models.advanced import handlers as _b200


class RCANB200Handler(_b200.RCANHandler):
    pass


class EDSRB200Handler(_b200.EDSRHandler):
    pass
