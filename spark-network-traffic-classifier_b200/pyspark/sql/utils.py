"""pyspark.sql.utils: the exception names user code catches."""
from . import AnalysisException  # noqa: F401
from ..ml.feature import IllegalArgumentException  # noqa: F401
