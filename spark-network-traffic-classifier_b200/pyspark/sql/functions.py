"""pyspark.sql.functions used by the reference scripts (kdd99.py:8,27; cicids17.py:8,27,30-35)."""
from . import Column, _RegexpReplace


def col(name):
    return Column(name)


column = col


def regexp_replace(str_col, pattern, replacement):
    """regexp_replace(column, java-regex, replacement) on a string column (applied to its dictionary)."""
    return _RegexpReplace(str_col, pattern, replacement)
