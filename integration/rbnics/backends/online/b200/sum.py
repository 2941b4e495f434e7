"""B200 online backend: sum() of a product() output (rbnics/backends/online/basic/sum.py:8-9)."""
from rbnics.backends.online.basic import sum as basic_sum
from rbnics.backends.online.b200.product import ProductOutput
from rbnics.utils.decorators import backend_for


@backend_for("b200", inputs=(ProductOutput, ))
def sum(product_output):
    return basic_sum(product_output)
