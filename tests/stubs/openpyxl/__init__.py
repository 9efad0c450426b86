"""Test stub of `openpyxl` (not installed here): just enough of Workbook / load_workbook for eval.py:76-95, persisted as
JSON so the integration test can read back what the unmodified eval.py wrote."""
import json
import os
import re


class _Cell:
    def __init__(self, sheet, key):
        self._sheet, self._key = sheet, key

    @property
    def value(self):
        return self._sheet.cells.get(self._key)

    @value.setter
    def value(self, v):
        self._sheet.cells[self._key] = v if isinstance(v, (str, int)) or v is None else float(v)


class _Sheet:
    def __init__(self, title, cells=None):
        self.title, self.cells = title, dict(cells or {})

    def __getitem__(self, key):
        assert re.fullmatch(r'[A-Z]+[0-9]+', key), key
        return _Cell(self, key)


class Workbook:
    def __init__(self):
        self._sheets = {'Sheet': _Sheet('Sheet')}

    @property
    def sheetnames(self):
        return list(self._sheets)

    def __getitem__(self, name):
        return self._sheets[name]

    def create_sheet(self, title=None):
        self._sheets[title] = _Sheet(title)
        return self._sheets[title]

    def save(self, file_name):
        os.makedirs(os.path.dirname(os.path.abspath(file_name)), exist_ok=True)
        with open(file_name, 'w') as fh:
            json.dump({k: s.cells for k, s in self._sheets.items()}, fh)


def load_workbook(file_name):
    if not os.path.exists(file_name):
        raise FileNotFoundError(file_name)
    wb = Workbook()
    with open(file_name) as fh:
        wb._sheets = {k: _Sheet(k, v) for k, v in json.load(fh).items()}
    return wb
