"""Model registry mirror (/root/reference/rumpy/shared_framework/models/__init__.py:7-35): handler classes are
discovered by AST-scanning `<package>/SISR/models/<category>/handlers.py` and registered under the lower-cased
class name minus 'Handler'; `define_model(name, **kwargs)` instantiates them."""
import ast
import os
from pydoc import locate

code_base_directory = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ml_tasks = ['SISR']
available_models = {}

for task in ml_tasks:
    model_dir = os.path.join(code_base_directory, task, 'models')
    if not os.path.isdir(model_dir):
        continue
    for category in [f.name for f in os.scandir(model_dir) if (f.is_dir() and '__' not in f.name)]:
        handler_file = os.path.join(model_dir, category, 'handlers.py')
        if not os.path.isfile(handler_file):
            continue
        tree = ast.parse(open(handler_file, 'r').read())
        for _class in [node.name for node in ast.walk(tree) if isinstance(node, ast.ClassDef)]:
            available_models[_class.split('Handler')[0].lower()] = \
                'rumpy_b200.%s.models.%s.handlers.%s' % (task, category, _class)


def define_model(name, **kwargs):
    return locate(available_models[name])(**kwargs)
